"""Normalizing-flow preconditioner: the reference's ``pocomc.flow.Flow`` API on sm_100a kernels.

Mirrors pocomc/flow.py (class Flow: __init__ 46-90, forward 99-114, inverse 116-132, log_prob
134-147, sample 149-163, fit 165-384).  The flow arithmetic that the reference delegates to zuko
(MAF / NSF over a masked MLP hyper-network) is implemented here:
  * inference (no grad): libpmc_b200's degree-ordered sweep kernels (csrc/flow_sweep.cu) --
    one sweep replaces zuko's D+1 hyper-network passes for the inverse;
  * training (grad enabled): an autograd graph over the same parameters (north_star: "PyTorch
    only for tensor containers and autograd on the flow").
Tensors may live on the host (like the reference's) or on the GPU; results come back on the
input's device.  There is no CPU compute path: without a CUDA device every call raises.
"""
from __future__ import annotations

import copy
import math
import time
import warnings
from typing import Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, config
from . import made_layout as ML
from . import tri_layout as TL
from .tools import torch_double_to_float

__all__ = ["Flow", "regularization_loss", "MaskedAutoregressiveFlow"]

PRESETS = {"maf3": (ML.KIND_AFFINE, 3), "maf6": (ML.KIND_AFFINE, 6), "maf12": (ML.KIND_AFFINE, 12),
           "nsf3": (ML.KIND_RQS, 3), "nsf6": (ML.KIND_RQS, 6), "nsf12": (ML.KIND_RQS, 12)}
_LOG_SLOPE = math.log(1e-3)


def _device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------------------
# differentiable univariate transforms (training path)
# ------------------------------------------------------------------------------------------
def _softclip(a, ls):
    return a / (1 + abs(a / ls))


def _affine_forward(x, phi):
    """zuko MonotonicAffineTransform: y = x exp(ls) + shift, ladj = ls."""
    ls = _softclip(phi[..., 1], _LOG_SLOPE)
    return x * ls.exp() + phi[..., 0], ls


def _rqs_forward(x, phi, bins=8, bound=5.0):
    """zuko MonotonicRQSTransform forward + log|dy/dx|."""
    w = _softclip(phi[..., :bins], _LOG_SLOPE / 2)
    h = _softclip(phi[..., bins:2 * bins], _LOG_SLOPE / 2)
    d = _softclip(phi[..., 2 * bins:], _LOG_SLOPE)
    hx = bound * (2 * torch.cumsum(F.pad(F.softmax(w, dim=-1), (1, 0)), dim=-1) - 1)
    hy = bound * (2 * torch.cumsum(F.pad(F.softmax(h, dim=-1), (1, 0)), dim=-1) - 1)
    dv = torch.exp(F.pad(d, (1, 1)))
    k = torch.searchsorted(hx.detach(), x.detach()[..., None].contiguous()).squeeze(-1) - 1
    mask = (k >= 0) & (k < bins)
    k = k % bins
    k01 = torch.stack((k, k + 1), dim=-1)
    x0, x1 = torch.gather(hx, -1, k01).unbind(-1)
    y0, y1 = torch.gather(hy, -1, k01).unbind(-1)
    d0, d1 = torch.gather(dv, -1, k01).unbind(-1)
    s = (y1 - y0) / (x1 - x0)
    z = mask * (x - x0) / (x1 - x0)
    den = s + (d0 + d1 - 2 * s) * z * (1 - z)
    y = y0 + (y1 - y0) * (s * z ** 2 + d0 * z * (1 - z)) / den
    jac = s ** 2 * (2 * s * z * (1 - z) + d0 * (1 - z) ** 2 + d1 * z ** 2) / den ** 2
    return torch.where(mask, y, x), torch.log(jac) * mask


class MaskedAutoregressiveFlow(nn.Module):
    """T masked-autoregressive transforms (affine = zuko MAF, rqs = zuko NSF bins=8) over a
    diagonal-normal base; the object the reference stores as ``Flow.flow``.

    All parameters live in ONE flat fp32 blob ``raw`` laid out in module order per transform
    (W0,b0,W1,b1,...,W_out,b_out; torch [out,in]); initialised layer by layer exactly like
    ``nn.Linear.reset_parameters`` so a given ``torch.manual_seed`` yields the same weights as
    constructing the zuko flow.
    """

    def __init__(self, features: int, hidden: int, n_layers: int, transforms: int, kind: int, bins: int = 8):
        super().__init__()
        self.layout = ML.build_layout(int(features), int(hidden), int(n_layers), int(transforms), int(kind), bins)
        lay = self.layout
        raw = torch.empty(lay.raw_numel, dtype=torch.float32)
        off = 0
        for _ in range(lay.n_transforms):
            shapes = lay.raw_sizes
            for i in range(0, len(shapes), 2):
                wshape, bshape = shapes[i], shapes[i + 1]
                w = raw[off:off + wshape[0] * wshape[1]].view(wshape)
                off += w.numel()
                b = raw[off:off + bshape[0]]
                off += b.numel()
                nn.init.kaiming_uniform_(w, a=math.sqrt(5))
                bound = 1 / math.sqrt(wshape[1]) if wshape[1] > 0 else 0
                nn.init.uniform_(b, -bound, bound)
        self.raw = nn.Parameter(raw)
        # kernel-side weight layout: the TMA-streamed consumption-order stream when the network fits the
        # stream kernel's shared-memory budget, else the degree-sorted slab layout of the v1 kernel
        self._pack_entry = "pmc_flow_pack"
        if ML.stream_supported(lay.n_dim, lay.n_hidden, lay.n_layers, lay.kind, lay.bins):
            klay = ML.build_stream(lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.kind, lay.bins)
            self.packed_numel = klay.numel
        else:
            klay = lay
            self.packed_numel = lay.packed_numel
        self.register_buffer("gather", torch.from_numpy(klay.gather.copy()), persistent=False)
        self.register_buffer("meta", torch.from_numpy(klay.meta.copy()), persistent=False)
        self._meta_host = np.ascontiguousarray(klay.meta)
        self._packed = None
        self._packed_key = None
        self._masks = None
        # tensor-core (tcgen05) image of the dense masked MLP for the forward direction
        self._tc = None
        if ML.tc_supported(lay.n_dim, lay.n_hidden, lay.kind):
            self._tc = ML.build_tc(lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.kind, lay.bins)
            self.register_buffer("tc_gather", torch.from_numpy(self._tc.gather.copy()), persistent=False)
            self._tc_meta_host = np.ascontiguousarray(self._tc.meta)
        self._tc_packed = None
        self._tc_key = None
        # tensor-core block-triangular sweep (csrc/flow_tri.cu): the inverse direction of affine flows
        self._tri = None
        if TL.tri_supported(lay.n_dim, lay.n_hidden, lay.n_layers, lay.kind):
            self._tri = TL.build_tri(lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.kind, lay.bins)
            self.register_buffer("tri_gather", torch.from_numpy(self._tri.gather.copy()), persistent=False)
            self.register_buffer("tri_meta", torch.from_numpy(self._tri.meta.copy()), persistent=False)
            self._tri_meta_host = np.ascontiguousarray(self._tri.meta)
        self._tri_packed = None
        self._tri_key = None
        self._tri_ws = None

    # -- the inner plugin seam: a ready ``zuko.flows.Flow`` (pocomc/flow.py:87-88) -------------------------------------
    @staticmethod
    def _zuko_linears(hyper):
        """(linear layer, wrapped in a residual block?) of a zuko MaskedMLP, in evaluation order"""
        out = []
        for layer in hyper.children():
            lin = getattr(layer, "f", layer)           # Residual(f) wraps the hidden -> hidden layers
            if hasattr(lin, "weight") and hasattr(lin, "mask"):
                out.append((lin, lin is not layer))
            elif type(layer).__name__ not in ("ReLU",):
                raise ValueError(f"unsupported hyper-network layer {type(layer).__name__}: the kernels are built for ReLU masked MLPs")
        return out

    @classmethod
    def from_zuko(cls, flow):
        """Adopt a user-built ``zuko.flows.MAF`` / ``NSF`` (what pocomc's ``Flow(n_dim, flow=<zuko flow>)`` accepts): read
        its structure, check that it is one the kernels implement -- alternating feature order, 3 residual hidden layers
        of equal width, ReLU, affine or 8..-bin spline heads, zuko's own hidden-unit degree assignment (the masks are
        compared entry by entry) -- and copy its parameters into the flat blob.  Anything else raises ValueError."""
        ts = getattr(flow, "transform", None)
        ts = list(ts) if isinstance(ts, (nn.ModuleList, list, tuple)) else ([ts] if ts is not None else [])
        if not ts or not all(hasattr(t, "hyper") and hasattr(t, "order") for t in ts):
            raise ValueError("not a zuko masked-autoregressive flow (no .transform[i].hyper / .order)")
        features = int(ts[0].order.numel())
        lin0 = cls._zuko_linears(ts[0].hyper)
        if len(lin0) != 4:
            raise ValueError(f"the kernels are built for 3 hidden layers, this flow has {len(lin0) - 1}")
        hidden = int(lin0[0][0].weight.shape[0])
        total = int(lin0[-1][0].weight.shape[0]) // features
        if total == 2:
            kind, bins = ML.KIND_AFFINE, 8
        elif total >= 5 and (total + 1) % 3 == 0:
            kind, bins = ML.KIND_RQS, (total + 1) // 3
        else:
            raise ValueError(f"{total} parameters per feature: neither an affine (2) nor a spline (3 bins - 1) head")
        if kind == ML.KIND_RQS and bins != 8:
            raise ValueError("spline kernels are built for bins = 8 (zuko.flows.NSF default, pocomc/flow.py:65-86)")
        self = cls(features, hidden, 3, len(ts), kind, bins)
        lay, chunks = self.layout, []
        for t, tr in enumerate(ts):
            want = np.arange(features) if t % 2 == 0 else features - 1 - np.arange(features)
            if not np.array_equal(tr.order.detach().cpu().numpy(), want) or int(getattr(tr, "passes", features)) != features:
                raise ValueError("feature orders must alternate (zuko randperm=False) with fully autoregressive transforms (passes = features)")
            lins = cls._zuko_linears(tr.hyper)
            if len(lins) != 4 or [r for _, r in lins] != [False, True, True, False]:
                raise ValueError("the kernels are built for hidden_features=[H]*3 with residual=True (pocomc/flow.py:55-86)")
            for (lin, _), mask, shape in zip(lins, ML.masks(lay, t), lay.raw_sizes[0::2]):
                if tuple(lin.weight.shape) != tuple(shape) or lin.bias is None:
                    raise ValueError(f"layer shape {tuple(lin.weight.shape)} does not match hidden_features=[{hidden}]*3")
                if not np.array_equal(lin.mask.detach().cpu().numpy().astype(bool), mask):
                    raise ValueError("hyper-network masks differ from zuko's MaskedMLP degree assignment: cannot adopt this flow")
                chunks += [lin.weight.detach().reshape(-1), lin.bias.detach().reshape(-1)]
        with torch.no_grad():
            self.raw.copy_(torch.cat([c.to(torch.float32).cpu() for c in chunks]))
        return self

    @torch.no_grad()
    def export_to(self, flow):
        """write the (trained) parameters back into the zuko module they were adopted from"""
        off, raw = 0, self.raw.detach().cpu()
        for tr in flow.transform:
            for lin, _ in self._zuko_linears(tr.hyper):
                for prm in (lin.weight, lin.bias):
                    prm.copy_(raw[off:off + prm.numel()].view_as(prm))
                    off += prm.numel()
        assert off == raw.numel()

    # -- plumbing ---------------------------------------------------------------------------
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_packed"], st["_packed_key"], st["_masks"] = None, None, None
        st["_tc_packed"], st["_tc_key"] = None, None
        st["_tri_packed"], st["_tri_key"], st["_tri_ws"] = None, None, None
        st.pop("_fit_engine", None)
        return st

    def _apply(self, fn, *a, **k):
        self._packed, self._packed_key, self._masks = None, None, None
        self._tc_packed, self._tc_key = None, None
        self._tri_packed, self._tri_key, self._tri_ws = None, None, None
        self.__dict__.pop("_fit_engine", None)
        return super()._apply(fn, *a, **k)

    def ensure_cuda(self):
        if not self.raw.is_cuda:
            self.to(_device())
        return self

    def transform_params(self, t: int):
        """[(W, b), ...] views of transform ``t`` (torch [out, in])."""
        lay, out, off = self.layout, [], t * self.layout.raw_tstride
        for i in range(0, len(lay.raw_sizes), 2):
            ws, bs = lay.raw_sizes[i], lay.raw_sizes[i + 1]
            w = self.raw[off:off + ws[0] * ws[1]].view(ws)
            off += ws[0] * ws[1]
            b = self.raw[off:off + bs[0]]
            off += bs[0]
            out.append((w, b))
        return out

    def packed(self) -> torch.Tensor:
        """degree-sorted slab copy of ``raw`` for the sweep kernels; rebuilt when raw changes."""
        self.ensure_cuda()
        key = (self.raw.data_ptr(), self.raw._version)
        if self._packed is None or self._packed_key != key:
            if self._packed is None or self._packed.device != self.raw.device:
                self._packed = torch.empty(self.packed_numel, dtype=torch.float32, device=self.raw.device)
            _lib.call(self._pack_entry, _lib.ptr(self.raw.detach()), _lib.ptr(self.gather), _lib.ptr(self._packed),
                      self.packed_numel)
            self._packed_key = key
        return self._packed

    # -- inference: tensor-core dense forward ------------------------------------------------
    def tc_available(self) -> bool:
        return self._tc is not None

    def packed_tc(self) -> torch.Tensor:
        """TF32 hi/lo weight image (mask folded in) for csrc/flow_tc.cu; rebuilt when raw changes."""
        self.ensure_cuda()
        key = (self.raw.data_ptr(), self.raw._version)
        if self._tc_packed is None or self._tc_key != key:
            if self._tc_packed is None or self._tc_packed.device != self.raw.device:
                self._tc_packed = torch.empty(self._tc.numel, dtype=torch.float32, device=self.raw.device)
            _lib.call("pmc_flow_tc_pack", _lib.ptr(self.raw.detach()), _lib.ptr(self.tc_gather), _lib.ptr(self._tc_packed),
                      self._tc.numel)
            self._tc_key = key
        return self._tc_packed

    @torch.no_grad()
    def forward_tc_into(self, src: torch.Tensor, out: torch.Tensor, ladj: torch.Tensor, passes: int = 3):
        """data -> latent on tcgen05 (one dense masked-MLP pass per transform): CUDA f32 src/out [N, D], ladj [N]."""
        if self._tc is None:
            raise ValueError("this flow has no tensor-core forward (affine transforms, H in {32, 64, 128}, D <= 48)")
        packed = self.packed_tc()
        if not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()):
            raise ValueError("forward_tc_into needs a contiguous CUDA float32 input")
        _lib.call("pmc_flow_forward_tc", _lib.ptr(packed), self._tc_meta_host.ctypes.data_as(_lib.C.c_void_p),
                  int(self._tc_meta_host.size), _lib.ptr(src), _lib.ptr(out), _lib.ptr(ladj), src.shape[0], int(passes))

    # -- inference: tensor-core block-triangular sweep ---------------------------------------
    def tri_available(self) -> bool:
        return self.__dict__.get("_tri") is not None

    def _use_tri(self, inverse: bool) -> bool:
        """inverse: the default path of every affine flow the layout covers; forward: flows the dense tcgen05 kernel
        (csrc/flow_tc.cu, H <= 128) does not cover -- the BASELINE widths 256 / 512 / 1024"""
        if not self.tri_available():
            return False
        if self.layout.kind != ML.KIND_AFFINE and self.layout.n_hidden < config.tri_rqs_min_hidden:
            return False                     # narrow spline flows: the fp32-FMA stream sweep is as fast (tests/nsf_bench.py)
        if inverse:
            return config.inverse_path == "tri"
        return self._tc is None and config.forward_path == "tc" and config.inverse_path == "tri"

    def packed_tri(self) -> torch.Tensor:
        """update slabs (TF32 hi/lo) + in-block fp32 slabs for csrc/flow_tri.cu; rebuilt when raw changes."""
        self.ensure_cuda()
        key = (self.raw.data_ptr(), self.raw._version)
        if self._tri_packed is None or self._tri_key != key:
            if self._tri_packed is None or self._tri_packed.device != self.raw.device:
                self._tri_packed = torch.empty(self._tri.numel, dtype=torch.float32, device=self.raw.device)
            _lib.call("pmc_flow_tc_pack", _lib.ptr(self.raw.detach()), _lib.ptr(self.tri_gather), _lib.ptr(self._tri_packed),
                      self._tri.numel)
            self._tri_key = key
        return self._tri_packed

    def _tri_workspace(self, n: int):
        """scratch area of the windowed sweep (flows wider than one tensor-memory window): one slab per resident CTA"""
        if self._tri.ws_floats == 0:
            return None, 0
        need = int(_lib.load().pmc_flow_sweep_tri_workspace(self._tri_meta_host.ctypes.data_as(_lib.C.c_void_p),
                                                            int(self._tri_meta_host.size), int(n)))
        if need < 0:
            _lib.check(2, "pmc_flow_sweep_tri_workspace")
        ws = self.__dict__.get("_tri_ws")
        if ws is None or ws.numel() < need or ws.device != self.raw.device:
            ws = torch.empty(need, dtype=torch.float32, device=self.raw.device)
            self._tri_ws = ws
        return ws, need

    def _tri_args(self, src, out, ladj, inverse, passes=None):
        packed = self.packed_tri()
        if not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()):
            raise ValueError("the tensor-core sweep needs a contiguous CUDA float32 input")
        ws, need = self._tri_workspace(src.shape[0])
        return (_lib.ptr(packed), self._tri_meta_host.ctypes.data_as(_lib.C.c_void_p), _lib.ptr(self.tri_meta),
                int(self._tri_meta_host.size), _lib.ptr(src), _lib.ptr(out), _lib.ptr(ladj), src.shape[0], 1 if inverse else 0,
                int(config.tri_passes if passes is None else passes), _lib.ptr(ws), int(need)), (packed, ws)

    @torch.no_grad()
    def sweep_tri_into(self, src: torch.Tensor, out: torch.Tensor, ladj: torch.Tensor, inverse: bool, passes=None):
        """latent -> data (or data -> latent) on tcgen05: CUDA f32 src/out [N, D], ladj [N]."""
        if not self.tri_available():
            raise ValueError("this flow has no tensor-core block-triangular sweep (tri_layout.tri_supported)")
        args, _ = self._tri_args(src, out, ladj, inverse, passes)
        _lib.call("pmc_flow_sweep_tri", *args)

    # -- inference: sweep kernels -----------------------------------------------------------
    @torch.no_grad()
    def sweep(self, v: torch.Tensor, inverse: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """v [N, D] f32 (any device) -> (out [N, D], ladj [N]) on v's device."""
        if v.dim() != 2 or v.shape[1] != self.layout.n_dim:
            raise ValueError(f"expected input of shape (n, {self.layout.n_dim}), got {tuple(v.shape)}")
        if not inverse and self._tc is not None and config.forward_path == "tc" and v.shape[0] >= config.tc_min_rows:
            src = v.detach().to(self.raw.device, torch.float32).contiguous()
            out = torch.empty_like(src)
            ladj = torch.empty(src.shape[0], dtype=torch.float32, device=src.device)
            self.forward_tc_into(src, out, ladj)
            return out.to(v.device), ladj.to(v.device)
        if self._use_tri(inverse):
            src = v.detach().to(self.raw.device, torch.float32).contiguous()
            out = torch.empty_like(src)
            ladj = torch.empty(src.shape[0], dtype=torch.float32, device=src.device)
            self.sweep_tri_into(src, out, ladj, inverse)
            return out.to(v.device), ladj.to(v.device)
        packed = self.packed()
        src = v.detach().to(self.raw.device, torch.float32).contiguous()
        out = torch.empty_like(src)
        ladj = torch.empty(src.shape[0], dtype=torch.float32, device=src.device)
        _lib.call("pmc_flow_sweep", _lib.ptr(packed), _lib.ptr(self.meta),
                  self._meta_host.ctypes.data_as(_lib.C.c_void_p), int(self._meta_host.size), _lib.ptr(src),
                  _lib.ptr(out), _lib.ptr(ladj), src.shape[0], 1 if inverse else 0)
        return out.to(v.device), ladj.to(v.device)

    @torch.no_grad()
    def sweep_into(self, src: torch.Tensor, out: torch.Tensor, ladj: torch.Tensor, inverse: bool):
        """Allocation-free variant for the MCMC loop: CUDA f32 src/out [N, D], ladj [N]."""
        if self._use_tri(inverse):
            return self.sweep_tri_into(src, out, ladj, inverse)
        packed = self.packed()
        if not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()):
            raise ValueError("sweep_into needs a contiguous CUDA float32 input")
        _lib.call("pmc_flow_sweep", _lib.ptr(packed), _lib.ptr(self.meta),
                  self._meta_host.ctypes.data_as(_lib.C.c_void_p), int(self._meta_host.size), _lib.ptr(src),
                  _lib.ptr(out), _lib.ptr(ladj), src.shape[0], 1 if inverse else 0)

    def bind_sweep(self, src: torch.Tensor, out: torch.Tensor, ladj: torch.Tensor, inverse: bool):
        """Pre-bound ``sweep_into`` for fixed buffers and fixed weights (the MCMC loop calls the flow with the same
        tensors every step); the returned callable keeps the packed weight image alive."""
        if self._use_tri(inverse):
            args, keep = self._tri_args(src, out, ladj, inverse)
            run = _lib.bind("pmc_flow_sweep_tri", *args)
            run.keep = (keep, self._tri_meta_host, self.tri_meta, src, out, ladj)
            return run
        packed = self.packed()
        if not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()):
            raise ValueError("bind_sweep needs a contiguous CUDA float32 input")
        run = _lib.bind("pmc_flow_sweep", _lib.ptr(packed), _lib.ptr(self.meta),
                        self._meta_host.ctypes.data_as(_lib.C.c_void_p), int(self._meta_host.size), _lib.ptr(src),
                        _lib.ptr(out), _lib.ptr(ladj), src.shape[0], 1 if inverse else 0)
        run.keep = (packed, self.meta, self._meta_host, src, out, ladj)
        return run

    def mark_dirty(self):
        """``raw`` was updated in place by a kernel torch does not see (csrc/train_ops.cu): drop the packed copies."""
        self._packed_key, self._tc_key, self._tri_key = None, None, None

    # -- training: autograd graph over the same parameters -----------------------------------
    def _flat_mask(self):
        """MADE masks of every transform laid out like ``raw`` (1 for biases): one multiply per step
        yields all masked weights (zuko MaskedLinear: F.linear(x, mask * weight, bias))."""
        if self._masks is None or self._masks.device != self.raw.device:
            lay, parts = self.layout, []
            for t in range(lay.n_transforms):
                for i, m in enumerate(ML.masks(lay, t)):
                    parts.append(np.asarray(m, np.float32).reshape(-1))
                    parts.append(np.ones(lay.raw_sizes[2 * i + 1][0], np.float32))
            self._masks = torch.from_numpy(np.concatenate(parts)).to(self.raw.device)
            assert self._masks.numel() == self.raw.numel()
        return self._masks

    def _views(self, flat: torch.Tensor, t: int):
        lay, out, off = self.layout, [], t * self.layout.raw_tstride
        for i in range(0, len(lay.raw_sizes), 2):
            ws, bs = lay.raw_sizes[i], lay.raw_sizes[i + 1]
            w = flat[off:off + ws[0] * ws[1]].view(ws)
            off += ws[0] * ws[1]
            b = flat[off:off + bs[0]]
            off += bs[0]
            out.append((w, b))
        return out

    def forward_autograd(self, x: torch.Tensor, raw: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """data -> latent with a graph (1 masked-MLP pass per transform, like zuko's forward).
        ``raw``: differentiate with respect to this leaf (a detached alias of the blob) instead of the
        module parameter."""
        self.ensure_cuda()
        lay = self.layout
        x = x.to(self.raw.device, torch.float32)
        ladj = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
        masked = (self.raw if raw is None else raw) * self._flat_mask()
        for t in range(lay.n_transforms):
            params = self._views(masked, t)
            h = x
            last = len(params) - 1
            for i, (w, b) in enumerate(params):
                y = F.linear(h, w, b)
                if i == 0:
                    h = torch.relu(y)
                elif i < last:
                    h = torch.relu(h + y)       # residual hidden block (oracle/zuko/nn.py)
                else:
                    h = y
            phi = h.unflatten(-1, (lay.n_dim, lay.total))
            x, l = _affine_forward(x, phi) if lay.kind == ML.KIND_AFFINE else _rqs_forward(x, phi, lay.bins)
            ladj = ladj + l.sum(dim=-1)
        return x, ladj

    # -- the lazy-flow protocol the reference expects from ``zuko.flows.Flow`` ---------------
    def forward(self, c=None):
        return _BoundFlow(self)


class _Transform:
    """``flow().transform`` : call_and_ladj / inv.call_and_ladj (flow.py:114,131)."""

    def __init__(self, module: MaskedAutoregressiveFlow, inverse=False):
        self.module, self._inverse = module, inverse

    @property
    def inv(self):
        return _Transform(self.module, not self._inverse)

    def call_and_ladj(self, x):
        m = self.module
        if not self._inverse and torch.is_grad_enabled() and (m.raw.requires_grad or x.requires_grad):
            z, ladj = m.forward_autograd(x)
            return z.to(x.device), ladj.to(x.device)
        return m.sweep(x, self._inverse)

    def __call__(self, x):
        return self.call_and_ladj(x)[0]


class _BoundFlow:
    """``flow()`` : NormalizingFlow(transform, DiagNormal(0, I))."""

    def __init__(self, module):
        self.module = module
        self.transform = _Transform(module)

    def log_prob(self, x):
        z, ladj = self.transform.call_and_ladj(x)
        return (-0.5 * z ** 2 - 0.5 * math.log(2 * math.pi)).sum(dim=-1) + ladj

    def rsample_and_log_prob(self, shape=()):
        n = int(np.prod(shape)) if len(shape) else 1
        d = self.module.layout.n_dim
        z = torch.randn((n, d), dtype=torch.float32)        # host global torch RNG, like zuko's base.rsample
        x, ladj = self.transform.inv.call_and_ladj(z)
        lp = (-0.5 * z ** 2 - 0.5 * math.log(2 * math.pi)).sum(dim=-1) - ladj
        return x.reshape(*shape, d), lp.reshape(*shape)

    def rsample(self, shape=()):
        return self.rsample_and_log_prob(shape)[0]


def epoch_batches(n: int, batch_size: int, shuffle: bool):
    """Index batches of one pass of ``DataLoader(TensorDataset(..n rows..), batch_size, shuffle)``
    with the same draws from the global torch generator (flow.py:251-257,301,331):
    the iterator's base seed, then (shuffle only) RandomSampler's private seed + randperm."""
    torch.empty((), dtype=torch.int64).random_()
    if shuffle:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        g = torch.Generator()
        g.manual_seed(seed)
        perm = torch.randperm(n, generator=g)
    else:
        perm = torch.arange(n)
    return [perm[i:i + batch_size] for i in range(0, n, batch_size)]


HY_LR, HY_BETA1, HY_BETA2, HY_EPS, HY_WD, HY_CLIP = range(6)     # csrc/train_ops.cu


class _PlateauLR:
    """torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min', threshold_mode='abs', cooldown=0) on a
    plain float (flow.py:271-277,361); the new rate is written to the device hyper-parameter block."""

    def __init__(self, lr, factor=0.2, patience=20, threshold=1e-4, min_lr=1e-6, eps=1e-8):
        self.lr, self.factor, self.patience, self.threshold, self.min_lr, self.eps = lr, factor, patience, threshold, min_lr, eps
        self.best, self.bad = math.inf, 0

    def step(self, metric) -> bool:
        if metric < self.best - self.threshold:
            self.best, self.bad = metric, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            self.bad = 0
            new = max(self.lr * self.factor, self.min_lr)
            if self.lr - new > self.eps:
                self.lr = new
                return True
        return False


class _FitEngine:
    """Device-resident state of ``Flow.fit``: training matrix, AdamW moments, hyper-parameters, the
    epoch's batch index table, and ONE CUDA graph per (batch size, weighted) holding a whole optimiser
    step -- batch gather, autograd forward/backward over the flat blob, gradient clipping + AdamW
    (csrc/train_ops.cu), loss accumulation.  A batch costs one graph launch; ragged last batches reuse
    the full-size graph with zero-weight padding rows."""
    MAX_BATCHES = 512
    _staging_free = None    # CUDA event: the last upload out of the shared pinned batch tables has completed

    def __init__(self, module: "MaskedAutoregressiveFlow"):
        self.module = module
        dev = module.raw.device
        n = module.raw.numel()
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.hyper = torch.zeros(6, dtype=torch.float64, device=dev)
        self.scratch = torch.empty(int(_lib.load().pmc_adamw_scratch_size()), dtype=torch.float64, device=dev)
        self.gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.acc = torch.zeros((), dtype=torch.float64, device=dev)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        self.x = None
        self.w = None
        self.tables = {}        # Bp -> (idx_all [MAX_BATCHES, Bp] int64, mask_all [MAX_BATCHES, Bp] f32)
        self.graphs = {}        # (Bp, weighted, train) -> CUDAGraph
        self.launches = 0
        # hand-written forward/backward (csrc/flow_train.cu) when the flow shape is built, else autograd
        lay = module.layout
        self.fused = config.fit_kernels == "fused" and ML.train_supported(lay.n_dim, lay.n_hidden, lay.kind)
        if self.fused:
            tl = ML.build_train(lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.kind, lay.bins)
            self.tl = tl
            self.tl_meta = np.ascontiguousarray(tl.meta)
            self.tl_gather = torch.from_numpy(tl.gather.copy()).to(dev)
            self.tl_wmap = torch.from_numpy(tl.wmap.copy()).to(dev)
            self.tl_tiles = torch.from_numpy(tl.tiles.copy()).to(dev)
            self.tl_packed = torch.empty(tl.numel, dtype=torch.float32, device=dev)
            # inverse of the gather map: where parameter i sits in the forward / backward weight images
            g = tl.gather.astype(np.int64)
            pos = np.nonzero(g >= 0)[0]
            order = np.argsort(g[pos], kind="stable")
            raw_idx, where = g[pos][order], pos[order]
            first = np.ones(len(raw_idx), bool)
            first[1:] = raw_idx[1:] != raw_idx[:-1]
            assert np.bincount(raw_idx, minlength=n).max() <= 2, "a parameter appears in more than two image slots"
            pos_a = np.full(n, -1, np.int32)
            pos_b = np.full(n, -1, np.int32)
            pos_a[raw_idx[first]] = where[first]
            pos_b[raw_idx[~first]] = where[~first]
            self.tl_pos_a = torch.from_numpy(pos_a).to(dev)
            self.tl_pos_b = torch.from_numpy(pos_b).to(dev)
            self.grad = torch.zeros(n, dtype=torch.float32, device=dev)      # masked entries stay 0
            self.fscratch = {}      # Bp -> (scratch floats, loss partials)
            self.eval_partials = {} # Bp -> loss partials of a whole validation epoch
        # every other flow (spline heads = the reference's default presets, H >= 512, D > 64): the layer-wise kernels of
        # csrc/flow_train_lw.cu -- masked-linear GEMMs + univariate-head kernels, forward and backward, one C call per batch
        self.layerwise = (not self.fused) and config.fit_kernels != "autograd"
        if self.layerwise:
            self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
            self.fscratch = {}

    def load(self, x: torch.Tensor, w):
        """copy the training matrix (already shuffled like flow.py:229-234) into the static buffers"""
        n, d = x.shape
        if self.x is None or self.x.shape[0] < n:
            cap = max(4096, 1 << (int(n) - 1).bit_length())
            self.x = torch.zeros((cap, d), dtype=torch.float32, device=self.module.raw.device)
            self.w = torch.zeros(cap, dtype=torch.float32, device=self.module.raw.device)
            self.graphs.clear()                                   # graphs hold the old buffers
        self.x[:n].copy_(x)
        if w is not None:
            self.w[:n].copy_(w)

    def reset_optimizer(self, lr, weight_decay, clip):
        self.m.zero_(); self.v.zero_(); self.step.zero_()
        self.hyper.copy_(torch.tensor([lr, 0.9, 0.999, 1e-8, weight_decay, clip if clip is not None else 0.0], dtype=torch.float64))

    def set_lr(self, lr):
        self.hyper[HY_LR:HY_LR + 1].fill_(lr)

    def _tables(self, B):
        if B not in self.tables:
            dev = self.module.raw.device
            self.tables[B] = (torch.zeros((self.MAX_BATCHES, B), dtype=torch.int64, device=dev),
                              torch.zeros((self.MAX_BATCHES, B), dtype=torch.float32, device=dev))
        return self.tables[B]

    def _body_fused(self, B, weighted, train):
        """fused forward + backward + weight gradients -> clip + AdamW (+ loss / cursor bookkeeping + image update):
        four launches per optimiser step"""
        mod = self.module
        idx_all, mask_all = self._tables(B)
        if B not in self.fscratch:
            nfl = int(_lib.load().pmc_flow_train_scratch_size(self.tl_meta.ctypes.data_as(_lib.C.c_void_p), B))
            self.fscratch[B] = (torch.empty(nfl, dtype=torch.float32, device=mod.raw.device),
                                torch.zeros(B // 8, dtype=torch.float64, device=mod.raw.device))
        scratch, partials = self.fscratch[B]
        if not train:
            _lib.call("pmc_flow_pack", _lib.ptr(mod.raw), _lib.ptr(self.tl_gather), _lib.ptr(self.tl_packed), self.tl.numel)
        _lib.call("pmc_flow_train_step", _lib.ptr(self.tl_packed), self.tl_meta.ctypes.data_as(_lib.C.c_void_p),
                  int(self.tl_meta.size), _lib.ptr(self.x), _lib.ptr(self.w) if weighted else None, _lib.ptr(idx_all),
                  _lib.ptr(mask_all), _lib.ptr(self.cursor), B, _lib.ptr(scratch), _lib.ptr(partials), None,
                  _lib.ptr(self.tl_tiles), _lib.ptr(self.tl_wmap), _lib.ptr(self.grad), 1 if train else 0)
        if train:
            # clip + AdamW, and in the same two launches: loss accumulation, batch cursor, and the updated weights
            # written straight into the training image (run_epoch packs it once per epoch; no pack between steps)
            _lib.call("pmc_adamw_clip_step_ex", _lib.ptr(mod.raw), _lib.ptr(self.grad), _lib.ptr(self.m), _lib.ptr(self.v),
                      mod.raw.numel(), _lib.ptr(self.hyper), _lib.ptr(self.step), _lib.ptr(self.scratch), _lib.ptr(self.gnorm),
                      _lib.ptr(partials), B // 8, _lib.ptr(self.acc), _lib.ptr(self.cursor),
                      _lib.ptr(self.tl_pos_a), _lib.ptr(self.tl_pos_b), _lib.ptr(self.tl_packed))
        else:
            self.acc += partials.sum()
            self.cursor += 1

    def _lw_buffers(self, B):
        if B not in self.fscratch:
            lay, lib = self.module.layout, _lib.load()
            nfl = int(lib.pmc_flow_train_lw_scratch_size(lay.n_dim, lay.n_hidden, lay.n_transforms, lay.total, self.module.raw.numel(), B))
            self.fscratch[B] = (torch.empty(nfl, dtype=torch.float32, device=self.module.raw.device),
                                torch.zeros(int(lib.pmc_flow_train_lw_partials(B)), dtype=torch.float64, device=self.module.raw.device))
        return self.fscratch[B]

    def _body_lw(self, B, weighted, train):
        """layer-wise forward + backward (csrc/flow_train_lw.cu) -> clip + AdamW with the loss / cursor bookkeeping folded in"""
        mod, lay = self.module, self.module.layout
        idx_all, mask_all = self._tables(B)
        scratch, partials = self._lw_buffers(B)
        _lib.call("pmc_flow_train_step_lw", _lib.ptr(mod.raw), _lib.ptr(mod._flat_mask()), lay.n_dim, lay.n_hidden, lay.n_transforms,
                  0 if lay.kind == ML.KIND_AFFINE else 1, mod.raw.numel(), _lib.ptr(self.x), _lib.ptr(self.w) if weighted else None,
                  _lib.ptr(idx_all), _lib.ptr(mask_all), _lib.ptr(self.cursor), B, _lib.ptr(scratch), _lib.ptr(partials),
                  _lib.ptr(self.grad), 1 if train else 0)
        if train:
            _lib.call("pmc_adamw_clip_step_ex", _lib.ptr(mod.raw), _lib.ptr(self.grad), _lib.ptr(self.m), _lib.ptr(self.v),
                      mod.raw.numel(), _lib.ptr(self.hyper), _lib.ptr(self.step), _lib.ptr(self.scratch), _lib.ptr(self.gnorm),
                      _lib.ptr(partials), partials.numel(), _lib.ptr(self.acc), _lib.ptr(self.cursor), None, None, None)
        else:
            self.acc += partials.sum()
            self.cursor += 1

    def loss_and_grad(self, rows: torch.Tensor, weighted: bool):
        """(loss, gradient blob) of one batch on the hand-written kernels, no parameter update (tests / diagnostics)."""
        assert self.fused or self.layerwise
        if self.layerwise:
            B = (len(rows) + 31) // 32 * 32
            idx_all, mask_all = self._tables(B)
            idx_all[0].zero_(); mask_all[0].zero_()
            idx_all[0, :len(rows)] = rows.to(idx_all.device)
            mask_all[0, :len(rows)] = 1.0
            self.cursor.zero_(); self.acc.zero_(); self.grad.zero_()
            scratch, partials = self._lw_buffers(B)
            mod, lay = self.module, self.module.layout
            _lib.call("pmc_flow_train_step_lw", _lib.ptr(mod.raw), _lib.ptr(mod._flat_mask()), lay.n_dim, lay.n_hidden, lay.n_transforms,
                      0 if lay.kind == ML.KIND_AFFINE else 1, mod.raw.numel(), _lib.ptr(self.x), _lib.ptr(self.w) if weighted else None,
                      _lib.ptr(idx_all), _lib.ptr(mask_all), _lib.ptr(self.cursor), B, _lib.ptr(scratch), _lib.ptr(partials),
                      _lib.ptr(self.grad), 1)
            return float(partials.sum().item()), self.grad.clone()
        B = (len(rows) + 31) // 32 * 32
        idx_all, mask_all = self._tables(B)
        idx_all[0].zero_(); mask_all[0].zero_()
        idx_all[0, :len(rows)] = rows.to(idx_all.device)
        mask_all[0, :len(rows)] = 1.0
        self.cursor.zero_(); self.acc.zero_()
        self.grad.zero_()
        mod = self.module
        if B not in self.fscratch:
            nfl = int(_lib.load().pmc_flow_train_scratch_size(self.tl_meta.ctypes.data_as(_lib.C.c_void_p), B))
            self.fscratch[B] = (torch.empty(nfl, dtype=torch.float32, device=mod.raw.device),
                                torch.zeros(B // 8, dtype=torch.float64, device=mod.raw.device))
        scratch, partials = self.fscratch[B]
        _lib.call("pmc_flow_pack", _lib.ptr(mod.raw), _lib.ptr(self.tl_gather), _lib.ptr(self.tl_packed), self.tl.numel)
        _lib.call("pmc_flow_train_step", _lib.ptr(self.tl_packed), self.tl_meta.ctypes.data_as(_lib.C.c_void_p),
                  int(self.tl_meta.size), _lib.ptr(self.x), _lib.ptr(self.w) if weighted else None, _lib.ptr(idx_all),
                  _lib.ptr(mask_all), _lib.ptr(self.cursor), B, _lib.ptr(scratch), _lib.ptr(partials), None,
                  _lib.ptr(self.tl_tiles), _lib.ptr(self.tl_wmap), _lib.ptr(self.grad), 1)
        return float(partials.sum().item()), self.grad.clone()

    def _body(self, B, weighted, optimise):
        mod = self.module
        idx_all, mask_all = self._tables(B)
        idx = idx_all.index_select(0, self.cursor).view(B)
        msk = mask_all.index_select(0, self.cursor).view(B)
        xb = self.x.index_select(0, idx)
        # a fresh leaf aliasing the blob: its autograd bookkeeping is born on the capturing stream and is
        # independent of whatever graph the user built on ``module.raw`` before
        leaf = mod.raw.detach().requires_grad_(True)
        z, ladj = mod.forward_autograd(xb, raw=leaf)
        lp = (-0.5 * z ** 2 - 0.5 * math.log(2 * math.pi)).sum(dim=-1) + ladj
        if weighted:
            wb = self.w.index_select(0, idx) * msk
            loss = (-lp * wb * 1000.0).sum() / wb.sum()           # flow.py:307-310
        else:
            loss = -(lp * msk).sum()                              # flow.py:305
        (g,) = torch.autograd.grad(loss, leaf)
        if optimise:
            _lib.call("pmc_adamw_clip_step", _lib.ptr(mod.raw), _lib.ptr(g), _lib.ptr(self.m), _lib.ptr(self.v), mod.raw.numel(),
                      _lib.ptr(self.hyper), _lib.ptr(self.step), _lib.ptr(self.scratch), _lib.ptr(self.gnorm))
            self.acc += loss.detach().double()
            self.cursor += 1

    STEPS_PER_GRAPH = 8     # consecutive optimiser steps captured in one graph (the batch cursor lives on the device)

    def graph(self, B, weighted, train=True, steps=1):
        key = (B, bool(weighted), bool(train), int(steps))
        if key not in self.graphs:
            self._tables(B)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if self.layerwise:
                    self.module._flat_mask()
                    self._lw_buffers(B)                          # allocate outside the capture
                elif not self.fused:
                    for _ in range(2):                           # warm-up without touching any state
                        self._body(B, weighted, optimise=False)
                else:
                    self.fscratch.get(B) or self._body_fused(B, weighted, train=False)   # allocate outside the capture
                    self.acc.zero_(); self.cursor.zero_()
            cur.wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(int(steps)):
                    if self.fused:
                        self._body_fused(B, weighted, train)
                    elif self.layerwise:
                        self._body_lw(B, weighted, train)
                    else:
                        self._body(B, weighted, optimise=True)
            self.graphs[key] = g
        return self.graphs[key]

    def run_epoch(self, batches, B, weighted, train=True, offset=0):
        """batches: list of host index tensors (each <= B rows) into the training matrix (shifted by
        ``offset``); returns the device scalar holding the summed batch losses.  ``train=False`` (hand-written
        kernels only) evaluates the loss without touching the parameters."""
        nb = len(batches)
        if nb > self.MAX_BATCHES:
            raise ValueError(f"more than {self.MAX_BATCHES} batches per epoch")
        nrows = B
        B = (B + 31) // 32 * 32                                  # the kernels work on 32-row tiles; extra rows are padding
        idx_all, mask_all = self._tables(B)
        hi = _lib.pinned("fit_idx", (nb, B), torch.int64)
        hm = _lib.pinned("fit_mask", (nb, B), torch.float32)
        # the staging buffers are shared by every epoch (and engine): the previous epoch's asynchronous upload must
        # have left them before the host refills them (a busy GPU delays that copy well past this point)
        ev = _FitEngine._staging_free
        if ev is not None:
            ev.synchronize()
        hi.zero_(); hm.zero_()
        for i, b in enumerate(batches):
            hi[i, :len(b)] = b + offset
            hm[i, :len(b)] = 1.0
        assert all(len(b) <= nrows for b in batches)
        g = None if (self.fused and not train) else self.graph(B, weighted, train)
        idx_all[:nb].copy_(hi, non_blocking=True)
        mask_all[:nb].copy_(hm, non_blocking=True)
        _FitEngine._staging_free = torch.cuda.Event()
        _FitEngine._staging_free.record()
        self.cursor.zero_()
        self.acc.zero_()
        if g is None:
            # validation pass: the batches are independent, so ONE launch covers the whole epoch (every batch keeps
            # its own weight normalisation, flow.py:307-310)
            mod = self.module
            if B not in self.eval_partials:
                self.eval_partials[B] = torch.zeros(self.MAX_BATCHES * (B // 8), dtype=torch.float64, device=mod.raw.device)
            partials = self.eval_partials[B]
            _lib.call("pmc_flow_pack", _lib.ptr(mod.raw), _lib.ptr(self.tl_gather), _lib.ptr(self.tl_packed), self.tl.numel)
            _lib.call("pmc_flow_eval_batches", _lib.ptr(self.tl_packed), self.tl_meta.ctypes.data_as(_lib.C.c_void_p),
                      int(self.tl_meta.size), _lib.ptr(self.x), _lib.ptr(self.w) if weighted else None, _lib.ptr(idx_all),
                      _lib.ptr(mask_all), _lib.ptr(self.cursor), B, nb, _lib.ptr(partials), None)
            self.acc += partials[:nb * (B // 8)].sum()
            self.launches += 1
            return self.acc
        if self.fused and train:       # the training image follows raw inside the step; bring it up to date once per epoch
            _lib.call("pmc_flow_pack", _lib.ptr(self.module.raw), _lib.ptr(self.tl_gather), _lib.ptr(self.tl_packed), self.tl.numel)
        k = self.STEPS_PER_GRAPH if (self.fused or self.layerwise) else 1
        if nb >= k > 1:
            gk = self.graph(B, weighted, train, steps=k)
            for _ in range(nb // k):
                gk.replay()
        else:
            k = nb + 1
        for _ in range(nb % k):
            g.replay()
        self.launches += nb
        if train:
            self.module.mark_dirty()
        return self.acc


class Flow:
    """
    Normalizing flow model (API of ``pocomc.Flow``).

    Parameters
    ----------
    n_dim : ``int``
        Number of dimensions of the distribution to be modeled.
    flow : ``str`` or ``MaskedAutoregressiveFlow``, optional
        One of ``maf3, maf6, maf12, nsf3, nsf6, nsf12`` (default ``nsf3``) or a ready module.
    """

    def __init__(self, n_dim, flow="nsf3"):
        self.n_dim = n_dim
        n_hidden = ML.hidden_width(int(n_dim))
        if isinstance(flow, str) and flow in PRESETS:
            kind, transforms = PRESETS[flow]
            self.flow = MaskedAutoregressiveFlow(n_dim, n_hidden, 3, transforms, kind)
        elif isinstance(flow, MaskedAutoregressiveFlow):
            self.flow = flow
        elif isinstance(flow, nn.Module) and hasattr(flow, "transform") and hasattr(flow, "base"):
            # the reference's inner plugin seam (flow.py:87-88): a ready zuko.flows.Flow.  Its parameters are copied into the
            # flat blob the kernels read; ``fit`` writes the trained values back into the user's module.
            self.flow = MaskedAutoregressiveFlow.from_zuko(flow)
            self.foreign = flow
        else:
            raise ValueError('Invalid flow type. Choose from: maf3, maf6, maf12, nsf3, nsf6, nsf12, '
                             'or provide a zuko.flows.Flow (MAF / NSF) object.')
        if torch.cuda.is_available():
            self.flow.ensure_cuda()

    @property
    def transform(self):
        return self.flow().transform

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """data -> latent: (u, log|du/dx|)   (flow.py:99-114)."""
        x = torch_double_to_float(x)
        return self.transform.call_and_ladj(x)

    __call__ = forward

    def inverse(self, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """latent -> data: (x, log|dx/du|)   (flow.py:116-132)."""
        u = torch_double_to_float(u)
        return self.transform.inv.call_and_ladj(u)

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """flow.py:134-147."""
        x = torch_double_to_float(x)
        return self.flow().log_prob(x)

    def sample(self, size: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """flow.py:149-163."""
        return self.flow().rsample_and_log_prob((size,))

    def fit(self, x, weights=None, validation_split=0.0, epochs=1000, batch_size=1000, patience=20,
            learning_rate=1e-3, weight_decay=0, laplace_scale=None, gaussian_scale=None, annealing=True,
            noise=None, shuffle=True, clip_grad_norm=1.0, verbose=0):
        """Weighted maximum-likelihood training, same control flow and RNG consumption as
        flow.py:165-384 (including its quirks): ``validation_split`` is the TRAIN
        fraction, losses are divided by the split size, best weights are restored only on early
        stop.  Data, parameters, optimiser state and the running losses stay on the GPU; the host
        synchronises once per epoch for the early-stopping test."""
        from torch.optim.lr_scheduler import ReduceLROnPlateau
        x = torch_double_to_float(x)
        module = self.flow.ensure_cuda()
        dev = module.raw.device
        n_samples, n_dim = x.shape
        if shuffle:
            rand_indx = torch.randperm(n_samples)
            x = x[rand_indx.to(x.device)]
            if weights is not None:
                weights = weights[rand_indx.to(weights.device)]
        x = x.to(dev)
        if weights is not None:
            weights = weights.to(dev, torch.float32)
        mean_min_dist = None
        if noise is not None:
            # reference quirk (flow.py:241-245): the mean is taken over the LAST row's distances
            mean_min_dist = torch.mean(torch.linalg.norm(x[-1] - x, dim=1))
        n_train = int(validation_split * n_samples) if validation_split > 0.0 else n_samples
        validation = validation_split > 0.0
        n_valid = n_samples - n_train

        use_graph = (config.fit_path == "graph" and noise is None and laplace_scale is None and gaussian_scale is None
                     and -(-n_train // max(int(batch_size), 1)) <= _FitEngine.MAX_BATCHES)
        engine = None
        if use_graph:
            engine = module.__dict__.get("_fit_engine")
            if engine is None:
                engine = module.__dict__["_fit_engine"] = _FitEngine(module)
            engine.load(x, weights)
            engine.reset_optimizer(learning_rate, weight_decay, clip_grad_norm)
            optimizer = None
            scheduler = _PlateauLR(learning_rate, factor=0.2, patience=patience, threshold=0.0001, min_lr=1e-6) if annealing else None
        else:
            optimizer = torch.optim.AdamW(module.parameters(), learning_rate, weight_decay=weight_decay)
            scheduler = None
            if annealing:
                scheduler = ReduceLROnPlateau(optimizer, mode='min', factor=0.2, patience=patience, threshold=0.0001,
                                              threshold_mode='abs', min_lr=1e-6)
        history = dict(loss=[], val_loss=[])
        monitor = 'val_loss' if validation else 'loss'
        best_epoch, best_loss = 0, np.inf
        best_model = module.raw.detach().clone()
        start = time.time()

        def batch_loss(idx, offset):
            idx = idx.to(dev) + offset
            xb = x[idx]
            if noise is not None:
                xb = xb + noise * mean_min_dist * torch.randn(xb.shape).to(dev)
            lp = module().log_prob(xb)
            if weights is None:
                loss = -lp.sum()
            else:
                wb = weights[idx]
                loss = (-lp * wb * 1000.0).sum() / wb.sum()
            if laplace_scale is not None or gaussian_scale is not None:
                loss = loss - regularization_loss(module, laplace_scale, gaussian_scale)
            return loss

        for epoch in range(epochs):
            module.train()
            if engine is not None:
                train_loss = engine.run_epoch(epoch_batches(n_train, batch_size, shuffle), int(batch_size), weights is not None)
            else:
                train_loss = torch.zeros((), dtype=torch.float64, device=dev)
                for idx in epoch_batches(n_train, batch_size, shuffle):
                    optimizer.zero_grad(set_to_none=True)
                    loss = batch_loss(idx, 0)
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_(module.parameters(), clip_grad_norm)
                    optimizer.step()
                    train_loss += loss.detach().double()
            val_loss = None
            if engine is not None and (engine.fused or engine.layerwise):
                train_loss = train_loss.clone()                  # engine.acc is reused by the validation pass
            if validation:
                module.eval()
                if engine is not None and (engine.fused or engine.layerwise):
                    val_loss = engine.run_epoch(epoch_batches(n_valid, batch_size, shuffle), int(batch_size), weights is not None,
                                                train=False, offset=n_train)
                else:
                    val_loss = torch.zeros((), dtype=torch.float64, device=dev)
                    with torch.no_grad():       # no graph needed: validation runs on the forward kernels
                        for idx in epoch_batches(n_valid, batch_size, shuffle):
                            val_loss += batch_loss(idx, n_train).double()
            train_loss = float(train_loss.item()) / n_train          # one sync per epoch
            history['loss'].append(train_loss)
            if validation:
                val_loss = float(val_loss.item()) / n_valid
                history['val_loss'].append(val_loss)
            if scheduler is not None:
                if engine is not None:
                    if scheduler.step(val_loss if validation else train_loss):
                        engine.set_lr(scheduler.lr)
                else:
                    scheduler.step(val_loss if validation else train_loss)
            if verbose > 1:
                if validation:
                    print('Epoch %3d/%3d, train loss: %5.2f, val loss: %5.2f' % (epoch + 1, epochs, train_loss, val_loss))
                else:
                    print('Epoch %3d/%3d, train loss: %5.2f' % (epoch + 1, epochs, train_loss))
            if history[monitor][-1] < best_loss:
                best_loss, best_epoch = history[monitor][-1], epoch
                best_model.copy_(module.raw.detach())
            if epoch - best_epoch >= int(1.5 * patience):
                with torch.no_grad():
                    module.raw.copy_(best_model)
                module.mark_dirty()
                if verbose > 0:
                    print('Finished early after %3d epochs' % best_epoch)
                    print('Best loss achieved %5.2f' % best_loss)
                break
        if verbose > 0:
            total = time.time() - start
            print()
            print('Time total:     %5.2f sec' % total)
            print('Time per epoch: %5.2f sec' % (total / epochs))
        if getattr(self, "foreign", None) is not None:
            module.export_to(self.foreign)        # the user's zuko module sees the trained parameters, like flow.py:268-370
        return history


def regularization_loss(model, laplace_scale=None, gaussian_scale=None):
    """L1 / L2 penalty on the hyper-network weights (flow.py:387-422).  ``model`` is the
    ``MaskedAutoregressiveFlow``; biases are excluded like the reference's name filter."""
    total_laplace, total_gaussian = 0.0, 0.0
    for t in range(model.layout.n_transforms):
        for w, _ in model.transform_params(t):
            if laplace_scale is not None:
                total_laplace = total_laplace + w.abs().sum()
            if gaussian_scale is not None:
                total_gaussian = total_gaussian + w.square().sum()
    total = 0.0
    if laplace_scale is not None:
        total = total - total_laplace / laplace_scale
    if gaussian_scale is not None:
        total = total - total_gaussian / (2.0 * gaussian_scale ** 2.0)
    return total
