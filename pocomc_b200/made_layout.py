"""Host-side layout of the MADE hyper-networks for the sm_100a sweep kernels.

The reference inverts an autoregressive transform with D full hyper-network passes plus one for
the log-determinant (zuko ``transform.inv.call_and_ladj``, reached from pocomc/flow.py:131 and
pocomc/mcmc.py:88).  Hidden unit ``h`` of every hidden layer has autoregressive degree
``(h mod (D-1)) + 1`` (zuko MaskedMLP pattern assignment), so a unit's activation is
final as soon as the inputs of lower order are.  We therefore sort hidden units by degree and
inputs/outputs by order; every mask becomes block lower-triangular and ONE degree-ordered sweep
computes each unit and each output exactly once.

This module builds, once per (D, H, L, T, kind):
  * ``meta``   int32 table describing the degree groups and the slab offsets, and
  * ``gather`` int32 map ``packed[i] = raw[gather[i]]`` (``-1`` -> 0.0) from the flat parameter
    blob (module order: per transform W0,b0,W1,b1,...,W_out,b_out, torch ``[out,in]`` layout)
    to the packed slab layout the kernels read.
The pack itself runs on the device (csrc/flow_sweep.cu: pmc_flow_pack).

Packed layout of one transform (float offsets relative to the transform's block):
  group g = 1..D-1 holds the sorted hidden units [gstart[g], gstart[g+1]) of degree g, padded
  to ``4*nchunk[g]`` columns.
  W0 slab g : [g rows (input orders 0..g-1)]      x [4*nchunk[g]]
  Wl slab g : [E_g rows (sorted units of degree<=g)] x [4*nchunk[g]]   (l = 1..L-1)
  Wo slab k : [E_k rows (sorted units of degree<=k)] x [TP]            (k = 0..D-1, E_0 = 0)
  biases    : b0/bl in padded-slot order [HP], b_out [D][TP] in order position.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache

import numpy as np

KIND_AFFINE, KIND_RQS = 0, 1
META_HEADER = 32          # int32 slots before the tables

# header slots
M_D, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE, M_HP, M_MAXCH = range(11)
M_OFF_GSTART, M_OFF_NCHUNK, M_OFF_SLOT, M_OFF_W0, M_OFF_WH, M_OFF_WO, M_OFF_B0, M_OFF_BH, M_OFF_BO, \
    M_RAW_TSTRIDE, M_BINS = range(11, 22)


def hidden_width(n_dim: int) -> int:
    """pocomc/flow.py:49-52."""
    n = 3 * n_dim
    p = 1 if n == 0 else 2 ** (n - 1).bit_length()
    return max(p, 32)


@dataclass(frozen=True)
class MadeLayout:
    n_dim: int
    n_hidden: int
    n_layers: int          # number of hidden layers L
    n_transforms: int
    kind: int
    bins: int
    total: int             # univariate parameters per dimension (2 affine, 3*bins-1 rqs)
    tp: int                # total padded to a multiple of 4
    raw_sizes: tuple       # per-transform raw tensor shapes in module order
    raw_tstride: int
    packed_tstride: int
    meta: np.ndarray       # int32
    gather: np.ndarray     # int32 [T * packed_tstride]
    hperm: np.ndarray      # sorted position -> original hidden unit
    degree: np.ndarray     # degree of sorted unit

    @property
    def raw_numel(self):
        return self.raw_tstride * self.n_transforms

    @property
    def packed_numel(self):
        return self.packed_tstride * self.n_transforms

    def order_perm(self, t):
        """input feature index with order position k in transform t (MAF alternates the order)."""
        k = np.arange(self.n_dim)
        return k if t % 2 == 0 else self.n_dim - 1 - k


@lru_cache(maxsize=None)
def build_layout(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8) -> MadeLayout:
    D, H, L, T = n_dim, n_hidden, n_layers, n_transforms
    if D < 2:
        # zuko: "The adjacency matrix leads to a null Jacobian." for a 1-D autoregressive net
        raise ValueError("The adjacency matrix leads to a null Jacobian.")
    total = 2 if kind == KIND_AFFINE else 3 * bins - 1
    tp = (total + 3) // 4 * 4
    ng = D - 1
    deg_orig = (np.arange(H) % ng) + 1
    hperm = np.argsort(deg_orig, kind="stable").astype(np.int64)
    degree = deg_orig[hperm]
    gstart = np.searchsorted(degree, np.arange(1, ng + 2), side="left").astype(np.int64)  # [ng+1]
    gsize = np.diff(gstart)
    nchunk = (gsize + 3) // 4
    slot = np.concatenate([[0], np.cumsum(4 * nchunk)]).astype(np.int64)     # padded slot start per group
    HP = int(slot[-1])
    E = gstart[1:]                                   # E_g for g = 1..ng  (units with degree <= g)
    Ek = np.concatenate([[0], E]).astype(np.int64)   # E_k for output order k = 0..D-1

    # raw (module-order) tensor offsets inside one transform
    O = D * total
    shapes = [(H, D), (H,)]
    for _ in range(L - 1):
        shapes += [(H, H), (H,)]
    shapes += [(O, H), (O,)]
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in shapes])]).astype(np.int64)
    raw_tstride = int(raw_off[-1])

    # packed offsets
    cur = 0
    off_w0 = np.zeros(ng, np.int64)
    for g in range(1, ng + 1):
        off_w0[g - 1] = cur
        cur += g * 4 * nchunk[g - 1]
    off_wh = np.zeros((max(L - 1, 1), ng), np.int64)
    for l in range(L - 1):
        for g in range(1, ng + 1):
            off_wh[l, g - 1] = cur
            cur += E[g - 1] * 4 * nchunk[g - 1]
    off_wo = np.zeros(D, np.int64)
    for k in range(D):
        off_wo[k] = cur
        cur += Ek[k] * tp
    off_b0 = cur
    cur += HP
    off_bh = np.zeros(max(L - 1, 1), np.int64)
    for l in range(L - 1):
        off_bh[l] = cur
        cur += HP
    off_bo = cur
    cur += D * tp
    packed_tstride = (cur + 3) // 4 * 4

    gather = np.full(T * packed_tstride, -1, np.int64)
    for t in range(T):
        base_p, base_r = t * packed_tstride, t * raw_tstride
        iperm = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
        w0 = base_r + raw_off[0]
        for g in range(1, ng + 1):
            wd = 4 * nchunk[g - 1]
            units = hperm[gstart[g - 1]:gstart[g]]
            rows = np.arange(g)
            blk = np.full((g, wd), -1, np.int64)
            blk[:, :len(units)] = w0 + units[None, :] * D + iperm[rows][:, None]
            gather[base_p + off_w0[g - 1]: base_p + off_w0[g - 1] + g * wd] = blk.reshape(-1)
            gather[base_p + off_b0 + slot[g - 1]: base_p + off_b0 + slot[g - 1] + len(units)] = \
                base_r + raw_off[1] + units
        for l in range(L - 1):
            wl, bl = base_r + raw_off[2 + 2 * l], base_r + raw_off[3 + 2 * l]
            for g in range(1, ng + 1):
                wd = 4 * nchunk[g - 1]
                units = hperm[gstart[g - 1]:gstart[g]]
                src = hperm[:E[g - 1]]
                blk = np.full((len(src), wd), -1, np.int64)
                blk[:, :len(units)] = wl + units[None, :] * H + src[:, None]
                o = base_p + off_wh[l, g - 1]
                gather[o:o + blk.size] = blk.reshape(-1)
                gather[base_p + off_bh[l] + slot[g - 1]: base_p + off_bh[l] + slot[g - 1] + len(units)] = bl + units
        wo, bo = base_r + raw_off[2 * L], base_r + raw_off[2 * L + 1]
        for k in range(D):
            feat = iperm[k]
            src = hperm[:Ek[k]]
            blk = np.full((len(src), tp), -1, np.int64)
            blk[:, :total] = wo + (feat * total + np.arange(total))[None, :] * H + src[:, None]
            o = base_p + off_wo[k]
            gather[o:o + blk.size] = blk.reshape(-1)
            o = base_p + off_bo + k * tp
            gather[o:o + total] = bo + feat * total + np.arange(total)

    tables = [gstart, nchunk, slot, off_w0, off_wh.reshape(-1), off_wo, off_bh]
    meta = np.zeros(META_HEADER, np.int64)
    meta[[M_D, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE, M_HP, M_MAXCH]] = \
        [D, H, L, T, kind, total, tp, ng, packed_tstride, HP, int(nchunk.max())]
    meta[M_RAW_TSTRIDE] = raw_tstride
    meta[M_BINS] = bins
    meta[M_OFF_B0] = off_b0
    meta[M_OFF_BO] = off_bo
    pos = META_HEADER
    for slot_id, tab in zip((M_OFF_GSTART, M_OFF_NCHUNK, M_OFF_SLOT, M_OFF_W0, M_OFF_WH, M_OFF_WO, M_OFF_BH), tables):
        meta[slot_id] = pos
        pos += len(tab)
    meta = np.concatenate([meta] + [np.asarray(t, np.int64) for t in tables])
    assert meta.max() < 2 ** 31 and gather.max() < 2 ** 31
    return MadeLayout(D, H, L, T, kind, bins, total, tp, tuple(shapes), raw_tstride, packed_tstride,
                      meta.astype(np.int32), gather.astype(np.int32), hperm, degree)


def masks(layout: MadeLayout, t: int):
    """Dense boolean masks (torch [out,in] layout, ORIGINAL unit order) of transform ``t`` --
    used by the training path (mask * weight) and by tests."""
    D, H, L, total = layout.n_dim, layout.n_hidden, layout.n_layers, layout.total
    order = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
    deg = (np.arange(H) % (D - 1)) + 1
    m0 = order[None, :] < deg[:, None]                       # [H, D]
    mh = deg[None, :] <= deg[:, None]                        # [H, H]
    mo = np.repeat(deg[None, :] <= order[:, None], total, axis=0)   # [D*total, H]
    return [m0] + [mh] * (L - 1) + [mo]
