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
# stream (v2) layout extras
M_VERSION, M_NCHUNKS, M_SLOT_FLOATS, M_OFF_CHUNKS = 22, 23, 24, 25
STREAM_CHUNK_FLOATS = 6144        # target size of one TMA bulk copy (24 KB)
STREAM_STAGES = 3                 # ring depth of the sweep kernel (keep in sync with flow_sweep.cu)
STREAM_SMEM_BUDGET = 200 * 1024   # ring + 8 particles of activations must fit below this


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


def useful_macs(layout: MadeLayout) -> int:
    """Multiply-accumulates per particle of one degree-ordered sweep over all transforms =
    nnz of the masks (the algorithmic minimum; the reference's inverse runs D+1 dense passes)."""
    D, H, L = layout.n_dim, layout.n_hidden, layout.n_layers
    deg = (np.arange(H) % (D - 1)) + 1
    per_t = int(deg.sum())                                              # input layer: unit of degree g sees g inputs
    le = np.array([(deg <= g).sum() for g in deg])
    per_t += (L - 1) * int(le.sum())                                    # hidden layers: units of degree <= own
    per_t += layout.total * int(sum((deg <= k).sum() for k in range(D)))   # outputs of order k see degree <= k
    return per_t * layout.n_transforms


@dataclass(frozen=True)
class StreamLayout:
    """Consumption-ordered weight stream for the TMA-fed sweep kernel (csrc/flow_sweep.cu, v2).

    One transform = D stages; stage k holds, in the order the kernel reads them,
      (every slab's row count is padded to a multiple of 16 with zero rows so the kernel's inner
      loop has no tail; the activation arrays it multiplies them with are zero-initialised)
      out hop  : TP/4 slabs [E_k rows][4] (outputs 4c..4c+3 of order position k) + bias [TP]
      group g=k+1 (if it has units), nch = ceil(size/4) chunks of 4 units:
        layer 0 : nch slabs [g rows][4] (input orders 0..g-1) + bias [4 nch]
        layer l : nch slabs [E_g rows][4] (sorted units of degree <= g of layer l-1) + bias [4 nch]
    Stages are grouped into chunks of ~24 KB; the kernel's producer warp streams the chunks
    through a shared-memory ring with cp.async.bulk + mbarriers.  The chunk table is the same for
    every transform: (k0, k1, float offset, float count)."""
    tstride: int
    slot_floats: int
    chunks: np.ndarray     # [n_chunks, 4] int64
    meta: np.ndarray       # int32
    gather: np.ndarray     # int32 [T * tstride]

    @property
    def numel(self):
        return int(self.gather.size)


def _pad16(n: int) -> int:
    return (int(n) + 15) // 16 * 16


def stream_supported(n_dim: int, n_hidden: int, n_layers: int, kind: int, bins: int = 8) -> bool:
    """The stream kernel needs ring + activations of >= 8 particles in shared memory."""
    total = 2 if kind == KIND_AFFINE else 3 * bins - 1
    tp = (total + 3) // 4 * 4
    wd = 4 * ((n_hidden + n_dim - 2) // max(n_dim - 1, 1) + 3) // 4 + 4
    stage_max = n_hidden * tp + tp + n_layers * (n_hidden * wd + wd)
    slot = max(stage_max, STREAM_CHUNK_FLOATS)
    return STREAM_STAGES * slot * 4 + 8 * (2 * n_dim + n_layers * n_hidden) * 4 + 4096 <= STREAM_SMEM_BUDGET


def _mma_slab(row_idx, col_idx_fn, n_cols):
    """One hop's weights in mma.sync m16n8k8 B-fragment order: [n-tile][k-step][lane][2] with
    lane = 4*g + t holding W[8*ks + t + 4*j][8*nt + g], j = 0, 1.  ``row_idx`` are the raw row keys
    (K of them, padded with zero rows to a multiple of 8); ``col_idx_fn(c, rows)`` gives the raw indices
    of column c for those rows (or None for a padding column)."""
    K = len(row_idx)
    K8 = (K + 7) // 8 * 8
    NT = (n_cols + 7) // 8
    dense = np.full((K8, NT * 8), -1, np.int64)
    for c in range(n_cols):
        col = col_idx_fn(c, row_idx)
        if col is not None:
            dense[:K, c] = col
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    out = np.empty((NT, K8 // 8, 32, 2), np.int64)
    for nt in range(NT):
        for ks in range(K8 // 8):
            for j in range(2):
                out[nt, ks, :, j] = dense[8 * ks + t + 4 * j, 8 * nt + g]
    return out.reshape(-1)


@lru_cache(maxsize=None)
def build_stream(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8,
                 variant: str = "ffma") -> StreamLayout:
    """variant "ffma": slabs [rows16][4] for the fp32-FMA stream kernel; "mma": B-fragment-ordered
    slabs for the warp-MMA (3xTF32 mma.sync) stream kernel."""
    mma = variant == "mma"
    lay = build_layout(n_dim, n_hidden, n_layers, n_transforms, kind, bins)
    D, H, L, T, total, tp = lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.total, lay.tp
    ng = D - 1
    hperm, degree = lay.hperm, lay.degree
    gstart = np.searchsorted(degree, np.arange(1, ng + 2), side="left").astype(np.int64)
    gsize = np.diff(gstart)
    nchunk = (gsize + 3) // 4
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(sh)) for sh in lay.raw_sizes])]).astype(np.int64)

    def transform_gather(t):
        """list of per-stage int64 gather arrays (indices into raw, -1 = zero)."""
        base_r = t * lay.raw_tstride
        iperm = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
        stages = []
        for k in range(D):
            parts = []
            feat = iperm[k]
            ek = int(gstart[k - 1 + 1]) if k >= 1 else 0          # units of degree <= k  (gstart[k] = #deg < k+1)
            src = hperm[:ek]
            wo, bo = base_r + raw_off[2 * L], base_r + raw_off[2 * L + 1]
            if mma:
                parts.append(_mma_slab(src, lambda c, rows: wo + (feat * total + c) * H + rows, total))
                ntp = (total + 7) // 8 * 8
                b = np.full(ntp, -1, np.int64)
                b[:total] = bo + feat * total + np.arange(total)
                parts.append(b)
            for c in range(0 if mma else tp // 4):
                blk = np.full((_pad16(ek), 4), -1, np.int64)
                for j in range(4):
                    o = 4 * c + j
                    if o < total:
                        blk[:ek, j] = wo + (feat * total + o) * H + src
                parts.append(blk.reshape(-1))
            if not mma:
                b = np.full(tp, -1, np.int64)
                b[:total] = bo + feat * total + np.arange(total)
                parts.append(b)
            g = k + 1
            if g <= ng and gsize[g - 1] > 0:
                units = hperm[gstart[g - 1]:gstart[g]]
                nch = int(nchunk[g - 1])
                upad = np.full(4 * nch, -1, np.int64)
                upad[:len(units)] = units
                eg = int(gstart[g])
                for l in range(L):
                    wl, bl = base_r + raw_off[2 * l], base_r + raw_off[2 * l + 1]
                    rows = iperm[np.arange(g)] if l == 0 else hperm[:eg]
                    width = D if l == 0 else H
                    if mma:
                        parts.append(_mma_slab(rows, lambda c, rr, wl=wl, width=width: wl + units[c] * width + rr, len(units)))
                        n8 = (len(units) + 7) // 8 * 8
                        bb = np.full(n8, -1, np.int64)
                        bb[:len(units)] = bl + units
                        parts.append(bb)
                        continue
                    for c in range(nch):
                        blk = np.full((_pad16(len(rows)), 4), -1, np.int64)
                        for j in range(4):
                            u = upad[4 * c + j]
                            if u >= 0:
                                blk[:len(rows), j] = wl + u * width + rows
                        parts.append(blk.reshape(-1))
                    parts.append(np.where(upad >= 0, bl + upad, -1))
            stages.append(np.concatenate(parts))
        return stages

    stages0 = transform_gather(0)
    sizes = np.array([len(a) for a in stages0], np.int64)
    assert np.all(sizes % 4 == 0)
    chunks, k0, acc, off = [], 0, 0, 0
    for k in range(D):
        if acc > 0 and acc + sizes[k] > STREAM_CHUNK_FLOATS:
            chunks.append((k0, k, off, acc))
            off += acc
            k0, acc = k, 0
        acc += int(sizes[k])
    chunks.append((k0, D, off, acc))
    chunks = np.asarray(chunks, np.int64)
    tstride = int(sizes.sum())
    slot_floats = int(chunks[:, 3].max())
    gather = np.concatenate([np.concatenate(transform_gather(t)) for t in range(T)])
    assert gather.size == T * tstride

    tables = [gstart, nchunk, chunks.reshape(-1)]
    meta = np.zeros(META_HEADER, np.int64)
    meta[[M_D, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE]] = [D, H, L, T, kind, total, tp, ng, tstride]
    meta[M_MAXCH] = int(nchunk.max())
    meta[M_RAW_TSTRIDE] = lay.raw_tstride
    meta[M_BINS] = bins
    meta[M_VERSION] = 3 if mma else 2
    meta[M_NCHUNKS] = len(chunks)
    meta[M_SLOT_FLOATS] = slot_floats
    pos = META_HEADER
    for slot_id, tab in zip((M_OFF_GSTART, M_OFF_NCHUNK, M_OFF_CHUNKS), tables):
        meta[slot_id] = pos
        pos += len(tab)
    meta = np.concatenate([meta] + [np.asarray(tb, np.int64) for tb in tables])
    assert meta.max() < 2 ** 31 and gather.max() < 2 ** 31
    return StreamLayout(tstride, slot_floats, chunks, meta.astype(np.int32), gather.astype(np.int32))


# ---------------------------------------------------------------------------------------------
# "tip" stream: bulk / tip split of every hop (csrc/flow_tip.cu, experimental)
# ---------------------------------------------------------------------------------------------
TIP_MAXCH = 2             # the kernel keeps one group's fresh activations in 4 * TIP_MAXCH registers per lane


def tip_supported(n_dim: int, n_hidden: int, n_layers: int, kind: int) -> bool:
    """Affine flows whose degree groups have at most 4 * TIP_MAXCH hidden units (H / (D - 1) <= 8: D >= 6 for
    the reference's H = max(next_pow2(3 D), 32))."""
    if kind != KIND_AFFINE or n_dim < 3 or n_layers < 1:
        return False
    return -(-n_hidden // (n_dim - 1)) <= 4 * TIP_MAXCH and stream_supported(n_dim, n_hidden, n_layers, kind)


@lru_cache(maxsize=None)
def build_stream_tip(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int = KIND_AFFINE,
                     bins: int = 8) -> StreamLayout:
    """Consumption-ordered weight stream for the bulk/tip sweep kernel.

    Every dot product of the degree-ordered sweep is split into the part over inputs that were finished one
    order position earlier (the BULK: no dependence on the value being computed right now, so its latency is
    hidden) and the part over the inputs born in the current position (the TIP: one degree group, <= 8 units).
    Stage k (order position k, feature iperm[k]; group g = k + 1 = sorted units [gstart[k], gstart[k+1])) holds,
    all in float4 units:
      out tip   : 4 nch(k-1) x (Wout[shift | scale][unit j of group k], 0, 0), then (b_shift, b_scale, 0, 0)
      bulk of g : layer 0   nch slabs [pad16(k) rows][4]          rows = inputs of order < k
                  layer l   nch slabs [pad16(gstart[k]) rows][4]   rows = sorted units of degree <= k
      out bulk  : (k + 1 < D) one slab [pad16(gstart[k]) rows][4] = (Wout[shift], Wout[scale], 0, 0) of feature k + 1
      tips of g : layer 0   per (chunk c, lane q): (bias, W0[unit, feature k], 0, 0)
                  layer l   per (chunk c, lane q): (bias, 0, 0, 0) + nch float4 of W_l[unit, units of group g]
    where unit = group unit 4 c + q (all zero for padding units).  meta: same header as build_stream, version 5."""
    if not tip_supported(n_dim, n_hidden, n_layers, kind):
        raise ValueError("flow shape not supported by the bulk/tip stream")
    lay = build_layout(n_dim, n_hidden, n_layers, n_transforms, kind, bins)
    D, H, L, T, total = lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.total
    ng = D - 1
    hperm, degree = lay.hperm, lay.degree
    gstart = np.searchsorted(degree, np.arange(1, ng + 2), side="left").astype(np.int64)    # gstart[i] = #units of degree <= i
    gsize = np.diff(gstart)
    nchunk = (gsize + 3) // 4
    assert nchunk.max() <= TIP_MAXCH and gsize.min() > 0
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(sh)) for sh in lay.raw_sizes])]).astype(np.int64)

    def padded_units(g):
        """raw unit ids of group g (1-based degree), padded with -1 to a multiple of 4"""
        u = hperm[gstart[g - 1]:gstart[g]]
        out = np.full(4 * int(nchunk[g - 1]), -1, np.int64)
        out[:len(u)] = u
        return out

    def transform_gather(t):
        base_r = t * lay.raw_tstride
        iperm = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
        wo, bo = base_r + raw_off[2 * L], base_r + raw_off[2 * L + 1]
        stages = []
        for k in range(D):
            parts = []
            feat = iperm[k]
            # ---- out tip (units of group k = degree k) + bias
            if k >= 1:
                for u in padded_units(k):
                    q = np.full(4, -1, np.int64)
                    if u >= 0:
                        q[:total] = wo + (feat * total + np.arange(total)) * H + u
                    parts.append(q)
            b = np.full(4, -1, np.int64)
            b[:total] = bo + feat * total + np.arange(total)
            parts.append(b)
            g = k + 1
            ek = int(gstart[k])                                   # sorted units of degree <= k
            if g <= ng:
                units = padded_units(g)
                nch = int(nchunk[g - 1])
                # ---- bulk of group g
                for l in range(L):
                    wl = base_r + raw_off[2 * l]
                    rows = iperm[np.arange(k)] if l == 0 else hperm[:ek]
                    width = D if l == 0 else H
                    for c in range(nch):
                        blk = np.full((_pad16(len(rows)), 4), -1, np.int64)
                        for j in range(4):
                            u = units[4 * c + j]
                            if u >= 0 and len(rows):
                                blk[:len(rows), j] = wl + u * width + rows
                        parts.append(blk.reshape(-1))
            if k + 1 < D:
                # ---- out bulk of feature k + 1 over last-layer units of degree <= k
                fnext = iperm[k + 1]
                blk = np.full((_pad16(ek), 4), -1, np.int64)
                for o in range(total):
                    blk[:ek, o] = wo + (fnext * total + o) * H + hperm[:ek]
                parts.append(blk.reshape(-1))
            if g <= ng:
                # ---- tips of group g
                for l in range(L):
                    wl, bl = base_r + raw_off[2 * l], base_r + raw_off[2 * l + 1]
                    for c in range(nch):
                        for q in range(4):
                            u = units[4 * c + q]
                            head = np.full(4, -1, np.int64)
                            if u >= 0:
                                head[0] = bl + u
                                if l == 0:
                                    head[1] = wl + u * D + feat
                            parts.append(head)
                            if l > 0:
                                tipw = np.full(4 * nch, -1, np.int64)
                                if u >= 0:
                                    ok = units >= 0
                                    tipw[ok] = wl + u * H + units[ok]
                                parts.append(tipw)
            stages.append(np.concatenate(parts))
        return stages

    stages0 = transform_gather(0)
    sizes = np.array([len(a) for a in stages0], np.int64)
    assert np.all(sizes % 4 == 0)
    chunks, k0, acc, off = [], 0, 0, 0
    for k in range(D):
        if acc > 0 and acc + sizes[k] > STREAM_CHUNK_FLOATS:
            chunks.append((k0, k, off, acc))
            off += acc
            k0, acc = k, 0
        acc += int(sizes[k])
    chunks.append((k0, D, off, acc))
    chunks = np.asarray(chunks, np.int64)
    tstride = int(sizes.sum())
    slot_floats = int(chunks[:, 3].max())
    gather = np.concatenate([np.concatenate(transform_gather(t)) for t in range(T)])
    assert gather.size == T * tstride
    tables = [gstart, nchunk, chunks.reshape(-1)]
    meta = np.zeros(META_HEADER, np.int64)
    meta[[M_D, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE]] = [D, H, L, T, kind, total, lay.tp, ng, tstride]
    meta[M_MAXCH] = int(nchunk.max())
    meta[M_RAW_TSTRIDE] = lay.raw_tstride
    meta[M_BINS] = bins
    meta[M_VERSION] = 5
    meta[M_NCHUNKS] = len(chunks)
    meta[M_SLOT_FLOATS] = slot_floats
    pos = META_HEADER
    for slot_id, tab in zip((M_OFF_GSTART, M_OFF_NCHUNK, M_OFF_CHUNKS), tables):
        meta[slot_id] = pos
        pos += len(tab)
    meta = np.concatenate([meta] + [np.asarray(tb, np.int64) for tb in tables])
    assert meta.max() < 2 ** 31 and gather.max() < 2 ** 31
    return StreamLayout(tstride, slot_floats, chunks, meta.astype(np.int32), gather.astype(np.int32))


# ---------------------------------------------------------------------------------------------
# tensor-core (tcgen05) layout of the DENSE masked MLP: Flow.forward / log_prob / training forward
# ---------------------------------------------------------------------------------------------
TC_D, TC_H, TC_L, TC_T, TC_KIND, TC_KX, TC_NOUT, TC_TSTRIDE, TC_BIAS_OFF, TC_NCHUNKS, TC_SLOT_BYTES, TC_VERSION, TC_LEN = range(13)
TC_KCHUNK = 32            # k extent of one streamed weight chunk (keep in sync with csrc/flow_tc.cu)
TC_BIAS_FLAG = 1 << 30    # gather entries copied unsplit (biases)


@dataclass(frozen=True)
class TcLayout:
    """Weight image of csrc/flow_tc.cu.  Per transform, for every linear layer l = 0..L (K_l = D padded
    to 8 for l = 0 else H; N_l = 2D padded to 16 for l = L else H) and every k-chunk of <= 32 columns:
    a TF32 ``hi`` image and a ``lo`` image, each in the no-swizzle K-major UMMA layout
    ``[k/4][N_l][4]`` (16-byte chunk of 4 consecutive k for output row n at (k/4)*N_l*16 + n*16), the
    MADE mask folded in (masked entries are 0).  The last chunk of a layer ends with the bias k-step
    ``[2][N_l][4]`` whose k = 0 / k = 1 entries are hi(b) / lo(b) (multiplied on the tensor core by a
    constant (1, 1, 0, ...) A block).  ``gather`` codes: >= 0 hi(raw[g]); -(g+2) lo(raw[g]);
    g | 2^30 plain copy; -1 zero."""
    tstride: int
    bias_off: int
    slot_bytes: int
    n_chunks: int
    meta: np.ndarray
    gather: np.ndarray

    @property
    def numel(self):
        return int(self.gather.size)


def tc_supported(n_dim: int, n_hidden: int, kind: int) -> bool:
    return kind == KIND_AFFINE and n_hidden in (32, 64, 128) and 2 <= n_dim <= 48 and 2 * n_dim <= 128


@lru_cache(maxsize=None)
def build_tc(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8) -> TcLayout:
    if not tc_supported(n_dim, n_hidden, kind):
        raise ValueError("flow not supported by the tensor-core forward kernel")
    lay = build_layout(n_dim, n_hidden, n_layers, n_transforms, kind, bins)
    D, H, L, T, total = lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.total
    Kx = (D + 7) // 8 * 8
    Nout = (D * total + 15) // 16 * 16
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in lay.raw_sizes])]).astype(np.int64)
    parts_all = []
    n_chunks = 0
    max_chunk = 0
    for t in range(T):
        base_r = t * lay.raw_tstride
        mks = masks(lay, t)
        parts = []
        for l in range(L + 1):
            K_true, K = (D, Kx) if l == 0 else (H, H)
            N_true, N = (D * total, Nout) if l == L else (H, H)
            w_off = base_r + raw_off[2 * l]
            idx = np.full((N, K), -1, np.int64)                      # raw index of W_l[n, k], -1 where masked / padding
            rr, cc = np.nonzero(mks[l])
            idx[rr, cc] = w_off + rr * K_true + cc
            for c0 in range(0, K, TC_KCHUNK):
                kc = min(TC_KCHUNK, K - c0)
                blk = idx[:, c0:c0 + kc].reshape(N, kc // 4, 4).transpose(1, 0, 2).reshape(-1)   # [k/4][N][4]
                parts.append(blk)                                    # hi image
                parts.append(np.where(blk >= 0, -(blk + 2), -1))     # lo image
                nbytes = 2 * kc * N * 4
                if c0 + kc >= K:                                     # bias k-step [2][N][4]: k = 0 -> hi(b), k = 1 -> lo(b)
                    bidx = base_r + raw_off[2 * l + 1] + np.arange(N_true)
                    bb = np.full((2, N, 4), -1, np.int64)
                    bb[0, :N_true, 0] = bidx
                    bb[0, :N_true, 1] = -(bidx + 2)
                    parts.append(bb.reshape(-1))
                    nbytes += N * 32
                max_chunk = max(max_chunk, nbytes)
                if t == 0:
                    n_chunks += 1
        w_floats = int(sum(len(a) for a in parts))
        parts_all.append(np.concatenate(parts))
    tstride = len(parts_all[0])
    assert tstride % 4 == 0 and w_floats % 4 == 0
    gather = np.concatenate(parts_all)
    slot_bytes = (max_chunk + 1023) // 1024 * 1024
    meta = np.zeros(TC_LEN, np.int64)
    meta[[TC_D, TC_H, TC_L, TC_T, TC_KIND, TC_KX, TC_NOUT, TC_TSTRIDE, TC_BIAS_OFF, TC_NCHUNKS, TC_SLOT_BYTES, TC_VERSION]] = \
        [D, H, L, T, kind, Kx, Nout, tstride, w_floats, n_chunks, slot_bytes, 100]
    assert np.abs(gather).max() < 2 ** 31
    return TcLayout(tstride, w_floats, slot_bytes, n_chunks, meta.astype(np.int32), gather.astype(np.int32))


# ---------------------------------------------------------------------------------------------
# fused training step (csrc/flow_train.cu): fp32 weight images for forward / input-gradient GEMMs and
# scatter maps for the weight-gradient GEMMs
# ---------------------------------------------------------------------------------------------
TR_D, TR_DP, TR_H, TR_L, TR_T, TR_NO, TR_TSTRIDE, TR_BIAS_OFF, TR_RAW_TSTRIDE, TR_MAP_TSTRIDE, TR_NTILES, TR_VERSION, TR_LEN = range(13)


@dataclass(frozen=True)
class TrainLayout:
    """Per transform (float offsets, every image a multiple of 4 floats):
      forward images  F_0 [D][H], F_1..F_{L-1} [H][H], F_o [H][No]      F_l[k][n] = (W_l * mask_l)[n][k],
                      each followed by its bias [N] (one bulk copy brings both)
      backward images B_o [No][H], B_{L-1}..B_1 [H][H], B_0 [H][Dp]      B_l[n][k] = (W_l * mask_l)[n][k]
    Output columns are permuted: column c = d + Dp*s holds shift (s = 0) / scale_raw (s = 1) of feature d
    (Dp = D rounded up to 32, No = 2 Dp), so one thread owns both parameters of a feature.
    ``gather``: packed[i] = raw[gather[i]] (-1 -> 0): the existing pmc_flow_pack kernel builds the image.
    ``wmap``: per transform, for every weight-gradient GEMM in the order layer 0..L, an int32 map
    [N_img][K_img] -> index into the flat gradient blob (-1: masked or padding), followed by the bias map
    [N_img].  ``tiles``: (t, l, n0, k0) of every 32x32 output tile of the weight-gradient GEMMs."""
    tstride: int
    bias_off: int
    map_tstride: int
    meta: np.ndarray
    gather: np.ndarray
    wmap: np.ndarray
    tiles: np.ndarray

    @property
    def numel(self):
        return int(self.gather.size)


def train_supported(n_dim: int, n_hidden: int, kind: int) -> bool:
    return kind == KIND_AFFINE and n_hidden in (32, 64, 128, 256) and 2 <= n_dim <= 64


@lru_cache(maxsize=None)
def build_train(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8) -> TrainLayout:
    if not train_supported(n_dim, n_hidden, kind):
        raise ValueError("flow not supported by the fused training kernels")
    lay = build_layout(n_dim, n_hidden, n_layers, n_transforms, kind, bins)
    D, H, L, T = lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms
    Dp = (D + 31) // 32 * 32
    No = 2 * Dp
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in lay.raw_sizes])]).astype(np.int64)
    # permuted output column c -> raw output row (2 d + s) or -1
    col2row = np.full(No, -1, np.int64)
    for s_ in range(2):
        col2row[s_ * Dp + np.arange(D)] = 2 * np.arange(D) + s_
    gathers, maps, tiles = [], [], []
    bias_off = None
    for t in range(T):
        base_r = t * lay.raw_tstride
        mks = masks(lay, t)
        idx = []                                           # per layer: raw index of (W*mask)[n][k] in IMAGE coords, -1 elsewhere
        for l in range(L + 1):
            K_true = D if l == 0 else H
            w_off = base_r + raw_off[2 * l]
            if l < L:
                full = np.where(mks[l], w_off + np.arange(H)[:, None] * K_true + np.arange(K_true)[None, :], -1)   # [H][K_true]
            else:
                full = np.full((No, H), -1, np.int64)
                ok = col2row >= 0
                rows = col2row[ok]
                full[ok] = np.where(mks[l][rows], w_off + rows[:, None] * H + np.arange(H)[None, :], -1)
            idx.append(full)
        bmaps = []
        for l in range(L + 1):
            b_off = base_r + raw_off[2 * l + 1]
            bmaps.append(b_off + np.arange(H) if l < L else np.where(col2row >= 0, b_off + col2row, -1))
        parts = []
        for l in range(L + 1):                             # forward images [K][N] followed by the bias [N]
            parts.append(idx[l].T.reshape(-1))
            parts.append(bmaps[l])
        for l in range(L, -1, -1):                         # backward images [N][K] (layer 0 padded to Dp columns)
            if l == 0:
                b0 = np.full((H, Dp), -1, np.int64)
                b0[:, :D] = idx[0]
                parts.append(b0.reshape(-1))
            else:
                parts.append(idx[l].reshape(-1))
        wf = int(sum(len(a) for a in parts))
        if bias_off is None:
            bias_off = wf
        gathers.append(np.concatenate(parts))
        mp = []
        for l in range(L + 1):
            mp.append(idx[l].reshape(-1))                  # [N_img][K_true]
            mp.append(bmaps[l])
            N_img, K_true = idx[l].shape
            for n0 in range(0, N_img, 32):
                for k0 in range(0, K_true, 32):
                    tiles.append((t, l, n0, k0))
        maps.append(np.concatenate(mp))
    tstride = len(gathers[0])
    assert tstride % 4 == 0 and bias_off % 4 == 0
    gather = np.concatenate(gathers)
    wmap = np.concatenate(maps)
    meta = np.zeros(TR_LEN, np.int64)
    meta[[TR_D, TR_DP, TR_H, TR_L, TR_T, TR_NO, TR_TSTRIDE, TR_BIAS_OFF, TR_RAW_TSTRIDE, TR_MAP_TSTRIDE, TR_NTILES, TR_VERSION]] = \
        [D, Dp, H, L, T, No, tstride, bias_off, lay.raw_tstride, len(maps[0]), len(tiles), 200]
    return TrainLayout(tstride, bias_off, len(maps[0]), meta.astype(np.int32), gather.astype(np.int32), wmap.astype(np.int32),
                       np.asarray(tiles, np.int32).reshape(-1, 4))


# ---------------------------------------------------------------------------------------------
# blocked sweep (csrc/flow_block.cu): degree blocks, dense part on the warp tensor path, triangular
# part as short fp32 dot products
# ---------------------------------------------------------------------------------------------
# The degree-ordered sweep is a (nonlinear) forward substitution: a unit of degree g needs every unit of
# degree <= g of the layer below.  Split the degrees into blocks.  Everything a block needs from EARLIER
# blocks is final before the block starts and is one dense product per layer ("phase A": [units of the
# block] x [all earlier units], mma.sync m16n8k8 with a 3xTF32 split, no dependency chain); what is left
# inside the block ("phase B") is the same hop-by-hop sweep with dot products over at most one block of
# units.  Phase A carries 75-90 % of the multiply-accumulates.
#
# The kernel is an interpreter over a per-transform PROGRAM of 8-int ops whose weights sit in the stream
# in exactly the order the ops consume them (one transform = one program pass; every transform has the
# same program, only the weights and the feature order differ):
#   OP_MMA   [0, src array (0: xs, l: act[l-1]), first src column, k-steps, dst array (l+1: act[l], L+1: ph),
#             first dst row, tiles (1..3 consecutive 16-row tiles sharing the B operand), flags]
#            flags FIRST (accumulators := bias fragments) / LAST (store to dst)
#            weights: [FIRST: tiles x 32 lanes x 4 bias]  then per k-step, per tile: [32 lanes x 4 hi][32 lanes x 4 lo]
#   OP_STEP  [1, k | ph_row << 16, out_rows | block_base << 16, first unit | units << 16, l0_col | l0_rows << 16,
#             hidden rows, 0, flags]: one order position k and the degree group g = k+1 behind it.  Only the NEW
#            values of the step travel through registers / shuffles (the critical chain x_k -> h_0 -> .. -> h_{L-1}
#            -> their share of output k+1); everything that was final before the step ("old" rows, read from shared
#            memory) is off that chain.  passes P = ceil(units / 4), slots = 4 P.  Weights, in order:
#              out_old [out_rows/4][2 params][4]          output k over the block's units of degree < k
#              l0_old  [l0_rows/4][slots][4], l0_new [slots]       layer 0: orders of the block before k, then order k
#              per layer l >= 1: old [hidden rows/4][slots][4] (block units of degree <= k), new [slots dst][slots src]
#              out_new [2 params][slots]   (flag NEXTOUT)  share of the group in output k+1, carried to the next step
#            flag BLOCKFIRST: nothing is carried into this step.
OP_MMA, OP_STEP = 0, 1
BF_FIRST, BF_LAST, BF_NEWCHUNK, BF_BLOCKFIRST, BF_NEXTOUT = 1, 2, 4, 8, 16
M_OFF_PROG, M_NOPS, M_HPB, M_SX, M_SO = 26, 27, 28, 29, 30
BLOCK_CHUNK_FLOATS = 4096         # target size of one bulk copy (16 KB)
BLOCK_STAGES = 4                  # ring depth (keep in sync with csrc/flow_block.cu)
BLOCK_MMA_FLOATS = 3328           # weights per OP_MMA piece (k-steps x tiles x 256 floats), below one chunk
BLOCK_MAX_STEPS = 8               # order positions per block (their 2 x 8 outputs are one 16-row MMA tile)
BLOCK_PLAIN = 1 << 30             # gather code: plain copy (pmc_flow_tc_pack: >= 0 hi, -(g+2) lo, -1 zero)
BLOCK_SMEM_BUDGET = 227 * 1024


@dataclass(frozen=True)
class BlockLayout:
    tstride: int
    slot_floats: int
    n_chunks: int
    n_ops: int
    hpb: int               # hidden units padded block-wise to multiples of 16
    sx: int                # row stride of xs [8 particles][sx]
    sh: int                # row stride of act[l] [8 particles][sh]
    so: int                # row stride of ph [8 particles][so]
    blocks: tuple          # (glo, ghi) degree range per block
    prog: np.ndarray       # [n_ops, 8] int32
    chunks: np.ndarray     # [n_chunks, 4] (0, 0, float offset, float count)
    meta: np.ndarray
    gather: np.ndarray

    @property
    def numel(self):
        return int(self.gather.size)

    def warp_floats(self, n_dim, n_layers):
        return 8 * n_dim + 8 * self.sx + n_layers * 8 * self.sh + 8 * self.so


def _degree_blocks(n_dim: int, n_hidden: int, block_units: int):
    """Consecutive degree ranges [glo, ghi): at most BLOCK_MAX_STEPS order positions and ~block_units units each."""
    ng = n_dim - 1
    gsize = np.bincount((np.arange(n_hidden) % ng) + 1, minlength=ng + 1)[1:]
    blocks, g = [], 1
    while g <= ng:
        ghi, cnt = g, 0
        while ghi <= ng and (cnt == 0 or cnt + gsize[ghi - 1] <= block_units + 16) and ghi - g < BLOCK_MAX_STEPS:
            cnt += int(gsize[ghi - 1])
            ghi += 1
        if ghi == ng + 1 and ghi - g == BLOCK_MAX_STEPS:       # the last block also owns the final output position
            ghi -= 1
        blocks.append((g, ghi))
        g = ghi
    return blocks, gsize


def block_supported(n_dim: int, n_hidden: int, n_layers: int, kind: int, block_units: int = 32) -> bool:
    """Affine transforms whose per-warp activations (8 particles) + ring fit shared memory with >= 3 warps."""
    if kind != KIND_AFFINE or n_dim < 3 or n_hidden < 2 * block_units:
        return False
    blocks, gsize = _degree_blocks(n_dim, n_hidden, block_units)
    if gsize.max() > 8:
        return False
    hpb = sum((int(((gsize[glo - 1:ghi - 1] + 3) // 4 * 4).sum()) + 15) // 16 * 16 for glo, ghi in blocks)
    per_warp = (8 * n_dim + 8 * (n_dim + 24) + n_layers * 8 * (hpb + 4) + 8 * 36) * 4
    n_ops = len(blocks) * (n_layers + 1) * (2 + hpb * 64 // BLOCK_MMA_FLOATS) + n_dim
    return BLOCK_STAGES * BLOCK_CHUNK_FLOATS * 4 + 3 * per_warp + 32 * n_ops + 4096 <= BLOCK_SMEM_BUDGET


@lru_cache(maxsize=None)
def build_block(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8,
                block_units: int = 32) -> BlockLayout:
    if kind != KIND_AFFINE:
        raise ValueError("the blocked sweep is built for affine (MAF) transforms")
    lay = build_layout(n_dim, n_hidden, n_layers, n_transforms, kind, bins)
    D, H, L, T, total = lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms, lay.total
    ng = D - 1
    hperm, degree = lay.hperm, lay.degree
    gstart = np.searchsorted(degree, np.arange(1, ng + 2), side="left").astype(np.int64)   # gstart[g-1] = first unit of degree g
    gsize = np.diff(gstart)
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(sh)) for sh in lay.raw_sizes])]).astype(np.int64)

    blocks, _ = _degree_blocks(D, H, block_units)
    nb = len(blocks)
    # padded unit index space: every degree group occupies 4 * passes slots (passes = ceil(size / 4) <= 2) so that a
    # group starts on a 16-byte boundary; every block is padded to whole 16-row MMA tiles
    npass = (gsize + 3) // 4
    assert npass.max() <= 2, "more than 8 hidden units per degree"
    u0_of = np.zeros(ng + 2, np.int64)              # padded index of the first unit of degree g
    pb = np.zeros(nb + 1, np.int64)
    for j, (glo, ghi) in enumerate(blocks):
        cur = int(pb[j])
        for g in range(glo, ghi):
            u0_of[g] = cur
            cur += 4 * int(npass[g - 1])
        pb[j + 1] = pb[j] + (cur - pb[j] + 15) // 16 * 16
    ps = np.diff(pb)
    hpb = int(pb[-1])
    pad2sorted = np.full(hpb, -1, np.int64)
    for g in range(1, ng + 1):
        pad2sorted[u0_of[g]:u0_of[g] + gsize[g - 1]] = np.arange(gstart[g - 1], gstart[g])
    pad2unit = np.where(pad2sorted >= 0, hperm[np.maximum(pad2sorted, 0)], -1)      # padded slot -> original hidden unit

    steps = [list(range(glo - 1, ghi - 1)) for glo, ghi in blocks]
    steps[-1].append(D - 1)            # the last output has no hidden group behind it
    max_steps = max(len(s) for s in steps)
    os_rows = (2 * max_steps + 15) // 16 * 16
    sx = (D + 7) // 8 * 8 + 8 + 4
    sh = hpb + 4
    so = os_rows + 4

    lane = np.arange(32)
    fr, fc = lane >> 2, lane & 3

    def a_fragments(dense):
        """dense [16, K8] raw indices (-1 = 0) -> per k-step hi image then lo image in m16n8k8 A-fragment order."""
        parts = []
        for ks in range(dense.shape[1] // 8):
            tile = dense[:, 8 * ks:8 * ks + 8]
            frag = np.stack([tile[fr, fc], tile[fr + 8, fc], tile[fr, fc + 4], tile[fr + 8, fc + 4]], axis=1).reshape(-1)
            parts.append(frag)
            parts.append(np.where(frag >= 0, -(frag + 2), -1))
        return parts

    def bias_fragment(b16):
        b16 = np.asarray(b16, np.int64)
        frag = np.stack([b16[fr], b16[fr], b16[fr + 8], b16[fr + 8]], axis=1).reshape(-1)
        return np.where(frag >= 0, frag | BLOCK_PLAIN, -1)

    def transform_program(t):
        """(ops [n,8] without chunk flags, list of per-op gather arrays) of transform t."""
        base_r = t * lay.raw_tstride
        iperm = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
        wl_off = [base_r + raw_off[2 * l] for l in range(L + 1)]
        bl_off = [base_r + raw_off[2 * l + 1] for l in range(L + 1)]
        ops, ws = [], []

        def emit_mma(src, kcols, dst, row0, denses, biases):
            """consecutive 16-row tiles sharing the B operand: bias + dense[16, kcols] . src[:, 0:kcols] each, split
            into pieces of at most BLOCK_MMA_FLOATS weights"""
            nt = len(denses)
            frs = [a_fragments(dn) if kcols > 0 else [] for dn in denses]
            nks = kcols // 8
            per = max(1, BLOCK_MMA_FLOATS // (256 * nt))
            pieces = [(k0, min(k0 + per, nks)) for k0 in range(0, nks, per)] if nks > 0 else [(0, 0)]
            for i, (k0, k1) in enumerate(pieces):
                flags = (BF_FIRST if i == 0 else 0) | (BF_LAST if i == len(pieces) - 1 else 0)
                w = [bias_fragment(b) for b in biases] if i == 0 else []
                for ks in range(k0, k1):
                    for tl in range(nt):
                        w += frs[tl][2 * ks:2 * ks + 2]
                ops.append([OP_MMA, src, 8 * k0, k1 - k0, dst, row0, nt, flags])
                ws.append(np.concatenate(w) if w else np.zeros(0, np.int64))

        def tile_groups(n_tiles):
            m = 0
            while m < n_tiles:
                take = min(3, n_tiles - m) if n_tiles - m != 4 else 2
                yield list(range(m, m + take))
                m += take

        def old_rows(row_units, dst_units, width, w_off, valid=None):
            """[rows/4][slots][4] block: weight of (dst unit slot, row) or -1; row_units / dst_units are raw unit / column
            keys (-1 = padding)"""
            rows, slots = len(row_units), len(dst_units)
            full = w_off + dst_units[None, :, None] * width + row_units.reshape(-1, 1, 4)
            ok = (dst_units >= 0)[None, :, None] & (row_units >= 0).reshape(-1, 1, 4)
            if valid is not None:
                ok = ok & valid.reshape(-1, 1, 4)
            return np.where(ok, full, -1).reshape(-1)

        for j, (glo, ghi) in enumerate(blocks):
            pbj = int(pb[j])
            # ---------------- phase A: contributions of everything before the block ----------------
            for l in range(L):
                kcols = (glo - 1 + 7) // 8 * 8 if l == 0 else pbj
                for grp in tile_groups(int(ps[j]) // 16):
                    denses, biases = [], []
                    for m in grp:
                        units = pad2unit[pbj + 16 * m: pbj + 16 * m + 16]
                        ok = units >= 0
                        dense = np.full((16, kcols), -1, np.int64)
                        if kcols > 0:
                            if l == 0:
                                c = np.arange(kcols)
                                cols = np.where(c < glo - 1, iperm[np.minimum(c, D - 1)], -1)
                                width = D
                            else:
                                cols = pad2unit[:kcols]
                                width = H
                            dense = np.where(ok[:, None] & (cols >= 0)[None, :], wl_off[l] + units[:, None] * width + cols[None, :], -1)
                        denses.append(dense)
                        biases.append(np.where(ok, bl_off[l] + units, -1))
                    emit_mma(0 if l == 0 else l, kcols, l + 1, pbj + 16 * grp[0], denses, biases)
            st = steps[j]
            for grp in tile_groups((2 * len(st) + 15) // 16):
                denses, biases = [], []
                for m in grp:
                    r = 16 * m + np.arange(16)
                    si, par = r // 2, r % 2
                    ok = si < len(st)
                    kk = np.asarray(st)[np.minimum(si, len(st) - 1)]
                    orow = iperm[kk] * total + par
                    dense = np.full((16, pbj), -1, np.int64)
                    if pbj > 0:
                        cols = pad2unit[:pbj]
                        dense = np.where(ok[:, None] & (cols >= 0)[None, :], wl_off[L] + orow[:, None] * H + cols[None, :], -1)
                    denses.append(dense)
                    biases.append(np.where(ok, bl_off[L] + orow, -1))
                emit_mma(L, pbj, L + 1, 16 * grp[0], denses, biases)
            # ---------------- phase B: the sweep inside the block ----------------
            for si, k in enumerate(st):
                feat = int(iperm[k])
                w = []
                # output hop of position k over the block's units of degree < k (degree k itself arrives in registers)
                out_old = int(u0_of[k]) - pbj if k >= glo else 0
                if out_old > 0:
                    rows = pad2unit[pbj:pbj + out_old]
                    w.append(old_rows(rows, feat * total + np.arange(2), H, wl_off[L]))
                g = k + 1
                u0 = cnt = l0_d0 = l0_r = lh = 0
                flags = (BF_BLOCKFIRST if si == 0 else 0) | (BF_NEXTOUT if si + 1 < len(st) else 0)
                if g <= ng and gsize[g - 1] > 0:
                    u0, cnt = int(u0_of[g]), int(gsize[g - 1])
                    P4 = 4 * int(npass[g - 1])
                    units = pad2unit[u0:u0 + P4]
                    lh = u0 - pbj
                    # layer 0: orders of the block before k ("old"), then order k itself
                    o_lo = glo - 1
                    l0_d0 = o_lo // 4 * 4
                    l0_r = (k - l0_d0 + 3) // 4 * 4
                    if l0_r > 0:
                        orders = l0_d0 + np.arange(l0_r)
                        okr = (orders >= o_lo) & (orders < k)
                        w.append(old_rows(np.where(okr, iperm[np.minimum(orders, D - 1)], -1), units, D, wl_off[0]))
                    w.append(np.where(units >= 0, wl_off[0] + units * D + feat, -1))
                    for l in range(1, L):
                        if lh > 0:
                            w.append(old_rows(pad2unit[pbj:u0], units, H, wl_off[l]))
                        nw = np.where((units >= 0)[:, None] & (units >= 0)[None, :], wl_off[l] + units[:, None] * H + units[None, :], -1)
                        w.append(nw.reshape(-1))                       # [dst slot][src slot]
                    if flags & BF_NEXTOUT:
                        nfeat = int(iperm[k + 1])
                        orow = nfeat * total + np.arange(2)
                        w.append(np.where((units >= 0)[None, :], wl_off[L] + orow[:, None] * H + units[None, :], -1).reshape(-1))
                assert max(k, 2 * si, out_old, pbj, u0, cnt, l0_d0, l0_r, lh) < 2 ** 15
                ops.append([OP_STEP, k | (2 * si) << 16, out_old | pbj << 16, u0 | cnt << 16, l0_d0 | l0_r << 16, lh, 0, flags])
                wcat = np.concatenate(w) if w else np.zeros(0, np.int64)
                ws.append(np.where(wcat >= 0, wcat | BLOCK_PLAIN, -1))
        return np.asarray(ops, np.int64), ws

    ops0, ws0 = transform_program(0)
    sizes = np.array([len(w) for w in ws0], np.int64)
    assert np.all(sizes % 4 == 0)
    chunks, off, acc = [], 0, 0
    for i in range(len(ops0)):
        if i == 0 or (acc > 0 and sizes[i] > 0 and acc + sizes[i] > BLOCK_CHUNK_FLOATS):
            if i > 0:
                chunks.append((0, 0, off, acc))
                off += acc
                acc = 0
            ops0[i, 7] |= BF_NEWCHUNK
        acc += int(sizes[i])
    chunks.append((0, 0, off, acc))
    chunks = np.asarray(chunks, np.int64)
    assert np.all(chunks[:, 3] > 0)
    tstride = int(sizes.sum())
    slot_floats = int(chunks[:, 3].max())
    gathers = []
    for t in range(T):
        ops_t, ws_t = transform_program(t)
        assert np.array_equal(ops_t[:, :7], ops0[:, :7]) and [len(w) for w in ws_t] == list(sizes)
        gathers.append(np.concatenate(ws_t))
    gather = np.concatenate(gathers)

    meta = np.zeros(META_HEADER, np.int64)
    meta[[M_D, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE]] = [D, H, L, T, kind, total, lay.tp, ng, tstride]
    meta[M_RAW_TSTRIDE] = lay.raw_tstride
    meta[M_BINS] = bins
    meta[M_VERSION] = 4
    meta[M_NCHUNKS] = len(chunks)
    meta[M_SLOT_FLOATS] = slot_floats
    meta[[M_NOPS, M_HPB, M_SX, M_SO]] = [len(ops0), hpb, sx, so]
    pos = META_HEADER
    meta[M_OFF_CHUNKS] = pos
    pos += chunks.size
    pos = (pos + 3) // 4 * 4                       # the program is read with 16-byte loads
    meta[M_OFF_PROG] = pos
    body = np.zeros(pos - META_HEADER + ops0.size, np.int64)
    body[:chunks.size] = chunks.reshape(-1)
    body[pos - META_HEADER:] = ops0.reshape(-1)
    meta = np.concatenate([meta, body])
    assert np.abs(gather[gather < BLOCK_PLAIN]).max() < 2 ** 29 and meta.max() < 2 ** 31
    return BlockLayout(tstride, slot_floats, len(chunks), len(ops0), hpb, sx, sh, so, tuple(blocks), ops0.astype(np.int32),
                       chunks, meta.astype(np.int32), gather.astype(np.int32))
