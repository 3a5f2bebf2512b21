"""ctypes binding of libpmc_b200.so (include/pmc_b200.h).  No CPU fallback: every compute entry
point raises if the library or a CUDA device is missing."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _build

_P, _I32, _I64, _F64, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_uint64


class PmcScaler(C.Structure):
    _fields_ = [("kind", _P), ("bc", _P), ("low", _P), ("high", _P), ("mu", _P), ("sigma", _P),
                ("log_sigma_sum", _F64), ("logit", _I32), ("scale", _I32)]


# name -> (restype, argtypes); mirrors include/pmc_b200.h one to one
SIGNATURES = {
    "pmc_last_error": (C.c_char_p, []),
    "pmc_version": (C.c_int, []),
    "pmc_device_info": (C.c_int, [_P, _P, _P]),
    "pmc_flow_pack": (C.c_int, [_P, _P, _P, _I64, _P]),
    "pmc_flow_sweep": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _I64, _I32, _P]),
    "pmc_flow_base_logprob": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "pmc_flow_tc_pack": (C.c_int, [_P, _P, _P, _I64, _P]),
    "pmc_flow_forward_tc": (C.c_int, [_P, _P, _I32, _P, _P, _P, _I64, _I32, _P]),
    "pmc_flow_sweep_tri_workspace": (_I64, [_P, _I32, _I64]),
    "pmc_flow_sweep_tri": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P]),
    "pmc_adamw_scratch_size": (_I64, []),
    "pmc_adamw_clip_step": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "pmc_adamw_clip_step_ex": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "pmc_flow_train_scratch_size": (_I64, [_P, _I64]),
    "pmc_flow_train_step": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _I32, _P]),
    "pmc_flow_eval_batches": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P]),
    "pmc_tpcn_propose": (C.c_int, [_I32, _P, _P, _P, _P, _F64, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "pmc_rwm_propose": (C.c_int, [_I32, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "pmc_scaler_inverse": (C.c_int, [_I32, _P, C.POINTER(PmcScaler), _P, _P, _P, _P, _I64, _I32, _P]),
    "pmc_scaler_inverse_prior": (C.c_int, [_I32, _P, C.POINTER(PmcScaler), _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "pmc_scaler_forward": (C.c_int, [_P, C.POINTER(PmcScaler), _P, _I64, _I32, _P]),
    "pmc_apply_bc": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P]),
    "pmc_mh_partials_size": (_I64, [_I64, _I32]),
    "pmc_mh_accept_update": (C.c_int, [_I32, _F64, _F64] + [_P] * 20 + [_I64, _I32, _P]),
    "pmc_mh_accept_finalize": (C.c_int, [_I32, _F64, _F64] + [_P] * 22 + [_I32, _I32, _I32, _I64, _I32, _P]),
    "pmc_mcmc_finalize": (C.c_int, [_I32, _P, _P, _I64, _P, _I32, _I32, _I32, _I64, _I32, _P]),
    "pmc_rng_fill": (C.c_int, [_U64, _U64, _I64, _F64, _P, _P, _P, _I64, _I32, _P]),
    "pmc_rng_fill_ctl": (C.c_int, [_U64, _P, _I64, _F64, _P, _P, _P, _I64, _I32, _P]),
    "pmc_ps_append": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I64, _P]),
    "pmc_ps_scratch_size": (_I64, [_I64]),
    "pmc_ps_reduce": (C.c_int, [_P, _P, _F64, _I32, _I64, _I64, _P, _P, _P]),
    "pmc_ps_weights": (C.c_int, [_P, _P, _F64, _I32, _I64, _P, _P, _P, _P]),
    "pmc_weight_stats": (C.c_int, [_P, _I64, _I64, _P, _P, _P]),
    "pmc_cumsum_f64": (C.c_int, [_P, _P, _I64, _P]),
    "pmc_resample_multinomial": (C.c_int, [_P, _P, _P, _I64, _I64, _P]),
    "pmc_resample_systematic": (C.c_int, [_P, _F64, _P, _I64, _I64, _P]),
    "pmc_gather_rows_f64": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "pmc_trim_threshold": (C.c_int, [_P, _I64, _F64, _I32, _P, _P, _P]),
    "pmc_trim_scratch_size": (_I64, [_I64]),
    "pmc_flow_train_lw_scratch_size": (_I64, [_I32, _I32, _I32, _I32, _I64, _I64]),
    "pmc_flow_train_lw_partials": (_I32, [_I64]),
    "pmc_flow_train_step_lw": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I64, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _I32, _P]),
    "pmc_geometry_scratch_size": (_I64, [_I64, _I32]),
    "pmc_weighted_colsums": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _P]),
    "pmc_weighted_scatter": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _P]),
    "pmc_mahalanobis": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P]),
    "pmc_student_weights": (C.c_int, [_P, _I64, _F64, _F64, _P, _P, _P, _P]),
    "pmc_comm_create": (C.c_int, [_I32, _I32, _I64, _P, _P]),
    "pmc_comm_connect": (C.c_int, [_P, _P, _P]),
    "pmc_comm_set_blocks": (C.c_int, [_P, _P, _P]),
    "pmc_comm_error": (C.c_int, [_P]),
    "pmc_comm_destroy": (C.c_int, [_P]),
    "pmc_mh_accept_finalize_p2p": (C.c_int, [_I32, _F64, _F64] + [_P] * 22 + [_I32, _I32, _I64, _I32, _P, _I64, _P]),
    "pmc_lse": (C.c_int, [_P, _I64, _P, _P, _P]),
    "pmc_lse_bootstrap": (C.c_int, [_P, _P, _I64, _I64, _P, _P]),
    "pmc_lse_bootstrap_rng": (C.c_int, [_P, _I64, _I64, _U64, _P, _P]),
    "pmc_loglike": (C.c_int, [_I32, _P, _P, _P, _F64, _F64, _P, _I64, _I32, _P]),
    "pmc_logprior": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "pmc_event_create": (C.c_int, [_P]),
    "pmc_event_destroy": (C.c_int, [_P]),
    "pmc_event_synchronize": (C.c_int, [_P]),
    "pmc_stream_synchronize": (C.c_int, [_P]),
    "pmc_memcpy_async": (C.c_int, [_P, _P, _I64, _P]),
    "pmc_download_rows": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P]),
}

_lib = None


def library_path() -> str:
    return _build.LIBPATH


def load():
    """dlopen the in-tree library (building it first if nvcc is around and it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIBPATH
    if not os.path.exists(path):
        _build.build()
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing: loud by design
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


_cuda_ok = False


def require_cuda():
    global _cuda_ok
    if _cuda_ok:
        return
    if not torch.cuda.is_available():
        raise RuntimeError("pocomc_b200: no CUDA device -- the hot path is CUDA-only (sm_100a); there is no CPU fallback")
    _cuda_ok = True


def device() -> torch.device:
    """the CUDA device this process computes on (one process per GPU)"""
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    """raw cudaStream_t of torch's current stream on the current device (cheap: no Stream object)."""
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


_pinned = {}


def pinned(tag: str, shape, dtype):
    """Process-wide cache of pinned host staging buffers (cudaHostAlloc costs milliseconds; the MCMC
    engine is rebuilt for every ``_mutate`` call).  One buffer per (tag, shape, dtype)."""
    key = (tag, tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,))), dtype)
    buf = _pinned.get(key)
    if buf is None:
        if len(_pinned) > 64:
            _pinned.clear()
        buf = torch.empty(key[1], dtype=dtype).pin_memory()
        _pinned[key] = buf
    return buf


def check(code: int, what: str = ""):
    if code != 0:
        msg = load().pmc_last_error()
        raise RuntimeError(f"libpmc_b200 {what} failed ({code}): {msg.decode() if msg else '?'}")


_entry_calls = [0]


def entry_calls() -> int:
    """Number of libpmc_b200 compute entry points invoked so far by this process (each launches at least one kernel):
    the counter behind bench.py's ``gpu_launches``."""
    return _entry_calls[0]


def call(name: str, *args):
    """Call an int-returning entry point on the current torch stream; raise on a non-zero code."""
    require_cuda()
    lib = load()
    _entry_calls[0] += 1
    check(getattr(lib, name)(*args, stream_ptr()), name)


def bind(name: str, *args, kernel: bool = True, stream: bool = True):
    """Pre-bound call for per-step hot loops: the argument tuple (fixed device pointers, sizes) is converted once;
    every invocation only looks up the current stream.  Same error behaviour as ``call``.  ``kernel=False``: the entry
    point launches no kernel (copies, events) and is left out of ``entry_calls``; ``stream=False``: it takes no stream."""
    require_cuda()
    fn = getattr(load(), name)
    get_stream, get_dev = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice

    counter = _entry_calls
    inc = 1 if kernel else 0

    if not stream:
        def run_plain():
            code = fn(*args)
            if code != 0:
                check(code, name)
        return run_plain

    def run():
        counter[0] += inc
        code = fn(*args, C.c_void_p(get_stream(get_dev())))
        if code != 0:
            check(code, name)
    return run


class Events:
    """A few CUDA events (timing disabled) for the chunked x' download of the MCMC step; process-wide cache by count."""
    _cache = {}

    def __init__(self, k: int):
        lib = load()
        self.handles = (C.c_void_p * k)()
        for i in range(k):
            h = C.c_void_p()
            check(lib.pmc_event_create(C.byref(h)), "pmc_event_create")
            self.handles[i] = h
        self.wait = [bind("pmc_event_synchronize", C.c_void_p(self.handles[i]), stream=False) for i in range(k)]

    @classmethod
    def get(cls, k: int) -> "Events":
        dev = torch.cuda.current_device()
        if (dev, k) not in cls._cache:
            cls._cache[(dev, k)] = cls(k)
        return cls._cache[(dev, k)]
