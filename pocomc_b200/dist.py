"""Particle sharding across the GPUs of one box (one process per GPU, torch.distributed/NCCL).

Rows (particles) are independent through the flow, proposal, reparameterisation and Metropolis
kernels given replicated flow weights / geometry / scaler, so rank r owns a contiguous block of
particles and the only per-MCMC-step exchange is the (D+4)-wide f64 reduction behind the sigma / mu
adaptation and the stop rule (mcmc.py:152,156,170).  Discrete decisions
downstream must not depend on the GPU count, so the reduction is NOT a ring all-reduce: every rank
contributes its fixed-size per-256-row block partials, they are all-gathered in rank (= global
particle) order and summed in that fixed order by every rank."""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.distributed as td

__all__ = ["init_from_env", "is_active", "world", "shard_range", "shard_counts", "gather_blocks", "allreduce_sum_det",
           "allreduce_sum_int", "broadcast_from_rank0", "barrier", "merge_weight_stats", "combine_weight_stats",
           "gather_history_scalars", "split_history_index", "peer_exchange", "peer_exchange_error"]


def init_from_env(backend: str = None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); no-op for one process."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or td.is_initialized():
        return world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    td.init_process_group(backend=backend)
    return world()


def is_active() -> bool:
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def world():
    """(rank, world_size)."""
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int, world_size: int, align: int = 1):
    """[start, stop) of the contiguous block of rows owned by ``rank``; block boundaries are
    multiples of ``align`` (256 = the accept kernel's rows-per-block keeps the block partials of a
    sharded run identical to the single-GPU ones)."""
    units = (n_total + align - 1) // align
    base, extra = divmod(units, world_size)
    start_u = rank * base + min(rank, extra)
    stop_u = start_u + base + (1 if rank < extra else 0)
    return min(start_u * align, n_total), min(stop_u * align, n_total)


def shard_counts(n_total: int, world_size: int, align: int = 1):
    """Rows owned by every rank under ``shard_range``."""
    return [b - a for a, b in (shard_range(n_total, r, world_size, align) for r in range(world_size))]


def _host_staged(t: torch.Tensor) -> bool:
    """gloo (the CPU / single-GPU test backend) has no all-gather for CUDA tensors: stage through the host."""
    return t.is_cuda and td.get_backend() == "gloo"


def gather_blocks(local: torch.Tensor, counts=None) -> torch.Tensor:
    """All-gather per-rank [b_r, w] rows (block partials, mutated particles) into [sum_r b_r, w] in
    rank order.  ``counts`` (rows per rank) may differ between ranks; equal counts take the
    single-collective path."""
    if not is_active():
        return local
    if _host_staged(local):
        return gather_blocks(local.cpu(), counts).to(local.device)
    if not local.is_cuda and td.get_backend() == "nccl":
        # NCCL moves device memory only: host arrays (history rows, index lists) are staged through this rank's GPU
        dev = torch.device("cuda", torch.cuda.current_device())
        return gather_blocks(local.to(dev), counts).cpu()
    ws = td.get_world_size()
    if counts is None:
        counts = [local.shape[0]] * ws
    if len(set(counts)) == 1:
        out = torch.empty((ws * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        td.all_gather_into_tensor(out, local.contiguous())
        return out
    # ragged: one equal-size collective over rows padded to the largest shard, padding dropped afterwards
    cmax = max(counts)
    padded = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((ws * cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    td.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * cmax:r * cmax + c] for r, c in enumerate(counts)], dim=0)


def allreduce_sum_det(values: torch.Tensor) -> torch.Tensor:
    """Sum of a small per-rank vector with a rank-ordered (GPU-count independent given the same
    per-rank values) association: gather then sum over the rank axis in order."""
    if not is_active():
        return values
    g = gather_blocks(values.reshape(1, -1))
    out = g[0].clone()
    for r in range(1, g.shape[0]):
        out += g[r]
    return out.reshape(values.shape)


def allreduce_sum_int(value: int) -> int:
    """Exact sum of one integer per rank (likelihood-call counters)."""
    if not is_active():
        return int(value)
    dev = "cpu" if td.get_backend() == "gloo" else torch.device("cuda", torch.cuda.current_device())
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    td.all_reduce(t)
    return int(t.item())


def broadcast_from_rank0(array: np.ndarray) -> np.ndarray:
    """rank 0's copy of a host array on every rank (seeds, prior samples: whatever every replica must agree on)"""
    if not is_active():
        return array
    dev = "cpu" if td.get_backend() == "gloo" else torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(np.ascontiguousarray(array)).to(dev)
    td.broadcast(t, src=0)
    return t.cpu().numpy()


def barrier():
    if is_active():
        td.barrier()


def merge_weight_stats(parts: np.ndarray) -> np.ndarray:
    """Merge per-shard (max logw, sum e, sum e^2) triples -- e = exp(logw - max) over the shard, what
    ``pmc_ps_reduce`` returns in out4[0:3] -- into the triple of the whole history, shard by shard in the
    given order with the rule of csrc/smc_ops.cu ``Lse3::merge`` (rescale the sums of the smaller maximum).
    Shards without finite weights (max = -inf) are skipped.  float64 throughout; pure host arithmetic."""
    parts = np.asarray(parts, dtype=np.float64).reshape(-1, 3)
    m, s1, s2 = -np.inf, 0.0, 0.0
    for om, o1, o2 in parts:
        if om == -np.inf:
            continue
        if m == -np.inf:
            m, s1, s2 = float(om), float(o1), float(o2)
        elif om <= m:
            c = math.exp(om - m)
            s1 += o1 * c
            s2 += o2 * c * c
        else:
            c = math.exp(m - om)
            s1 = s1 * c + o1
            s2 = s2 * c * c + o2
            m = float(om)
    return np.array([m, s1, s2], dtype=np.float64)


def combine_weight_stats(local3: torch.Tensor) -> torch.Tensor:
    """Sharded history ("every bisection probe"): every rank reduces its own slice of the
    history to (max, sum e, sum e^2); the triples are all-gathered and merged in rank order, so every rank
    holds the same global triple (hence the same ESS = s1^2 / s2 and log Z = max + log s1 - log(T N)) and
    takes the same branch of the bisection.  One 24-byte all-gather per probe."""
    if not is_active():
        return local3
    g = gather_blocks(local3.reshape(1, 3).to(torch.float64))
    merged = merge_weight_stats(g.detach().cpu().numpy())
    return torch.from_numpy(merged).to(local3.device)


def gather_history_scalars(local: torch.Tensor, counts) -> torch.Tensor:
    """History sharded by particle: every rank holds ``local [T, n_r]`` (one scalar per stored particle of its
    block, e.g. the persistent-sampling log-weights); returns the full ``[T, N]`` array in global particle
    order on every rank.  The bit-exact weight trimming (tools.trim_weights: global percentiles with numpy's
    linear interpolation) runs redundantly on this array -- 8 bytes per history element instead of the
    (2 D + 3) x 8 of the particle rows, which stay sharded."""
    if not is_active():
        return local
    rows = gather_blocks(local.t().contiguous(), list(counts))        # [N, T], rank blocks = particle blocks
    return rows.t().contiguous()


def split_history_index(idx: np.ndarray, n_total: int, counts) -> tuple:
    """Global flat history indices ``i = t * N + j`` (iteration t, particle j; what the trim / resampling
    kernels return) -> (owner rank of every index, flat index ``t * n_r + (j - lo_r)`` into that rank's
    ``[T, n_r]`` shard).  Pure index arithmetic, identical on every rank."""
    idx = np.asarray(idx, dtype=np.int64)
    counts = np.asarray(list(counts), dtype=np.int64)
    if counts.sum() != n_total:
        raise ValueError("shard counts do not add up to the number of particles")
    starts = np.concatenate([[0], np.cumsum(counts)])
    t, j = np.divmod(idx, n_total)
    owner = np.searchsorted(starts, j, side="right") - 1
    local = t * counts[owner] + (j - starts[owner])
    return owner, local


# ---------------------------------------------------------------------------------------------
# peer-memory exchange context (csrc/mcmc_ops.cu: pmc_comm_*): the per-MCMC-step reduction of a sharded run as stores into
# peer memory over NVLink from inside the accept kernel instead of an NCCL all-gather between two launches
# ---------------------------------------------------------------------------------------------
_peer = dict(ptr=None, capacity=0, blocks=None, unavailable=False)


def _all_ok(ok: bool) -> bool:
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
    td.all_reduce(t, op=td.ReduceOp.MIN)
    return bool(t.item())


def peer_exchange(block_counts, width: int):
    """The process-wide exchange context sized for ``sum(block_counts) * width`` doubles, or None when the ranks do not
    own distinct GPUs of one NCCL job (the gloo test set-up: both ranks on one device) or peer memory cannot be opened --
    the caller then takes the all-gather path.  Collective: every rank must call it with the same arguments."""
    from . import _lib, config
    if not (is_active() and td.get_backend() == "nccl" and config.p2p_exchange) or _peer["unavailable"]:
        return None
    rank, ws = world()
    if ws > 8:
        return None
    off = np.concatenate([[0], np.cumsum(np.asarray(block_counts, dtype=np.int64))]).astype(np.int32)
    need = int(off[-1]) * int(width)
    lib, C = _lib.load(), _lib.C
    if _peer["ptr"] is not None and _peer["capacity"] >= need:
        if _peer["blocks"] != tuple(off.tolist()):
            _lib.check(lib.pmc_comm_set_blocks(_peer["ptr"], off.ctypes.data_as(C.c_void_p), _lib.stream_ptr()), "pmc_comm_set_blocks")
            _peer["blocks"] = tuple(off.tolist())
            td.barrier()                              # nobody publishes into a table a peer has not switched yet
        return _peer["ptr"]
    dev = torch.device("cuda", torch.cuda.current_device())
    ids = torch.zeros(ws, dtype=torch.int64, device=dev)
    ids[rank] = torch.cuda.current_device() + 1
    td.all_reduce(ids)
    if len(set(ids.tolist())) != ws:                  # two ranks share a GPU
        _peer["unavailable"] = True
        return None
    if _peer["ptr"] is not None:
        torch.cuda.synchronize()
        td.barrier()
        lib.pmc_comm_destroy(_peer["ptr"])
        _peer.update(ptr=None, capacity=0, blocks=None)
    capacity = max(need * 2, 1 << 16)
    ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
    ok = lib.pmc_comm_create(rank, ws, capacity, C.byref(ptr), handle) == 0
    if not _all_ok(ok):
        _peer["unavailable"] = True
        return None
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    handles = torch.empty(ws * 64, dtype=torch.uint8, device=dev)
    td.all_gather_into_tensor(handles, mine)
    hbuf = (C.c_ubyte * (ws * 64))(*handles.cpu().tolist())
    ok = lib.pmc_comm_connect(ptr, hbuf, off.ctypes.data_as(C.c_void_p)) == 0
    if not _all_ok(ok):
        lib.pmc_comm_destroy(ptr)
        _peer["unavailable"] = True
        return None
    _peer.update(ptr=ptr, capacity=capacity, blocks=tuple(off.tolist()))
    td.barrier()
    return ptr


def peer_exchange_error() -> bool:
    """True when a kernel gave up waiting for a peer (the run is then stopped by the controller's stop flag)."""
    from . import _lib
    return _peer["ptr"] is not None and _lib.load().pmc_comm_error(_peer["ptr"]) == 1
