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


@lru_cache(maxsize=None)
def build_stream(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8) -> StreamLayout:
    """slabs [rows16][4] in consumption order for the fp32-FMA stream kernel (csrc/flow_sweep.cu)."""
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
            for c in range(tp // 4):
                blk = np.full((_pad16(ek), 4), -1, np.int64)
                for j in range(4):
                    o = 4 * c + j
                    if o < total:
                        blk[:ek, j] = wo + (feat * total + o) * H + src
                parts.append(blk.reshape(-1))
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
    meta[M_VERSION] = 2
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
# block-triangular sweep on tcgen05 (csrc/flow_tri.cu): Flow.inverse (and forward) of affine flows
# ---------------------------------------------------------------------------------------------
# The degree-ordered sweep is a nonlinear forward substitution.  Order positions (stages) are cut into blocks of
# TRI_G; the hidden units born in a block (degree groups k0+1 .. k0+8) form one K-slab.  Everything a block needs
# from EARLIER blocks is dense: after a block is finished, its activations (a [128 particles x K] A tile per layer,
# hi/lo TF32 images in shared memory) update the pre-activation accumulators of ALL later units with one
# tcgen05.mma group per layer (right-looking; accumulators live in tensor memory, one column per unit slot and per
# output).  What is left inside a block -- the block-triangular dependencies between its own TRI_G groups -- runs as fp32
# FMAs with one thread per particle and the block's activations in registers.  TRI_G = 4 keeps the fully unrolled
# in-block code of one block shape at ~7 KB, inside the SM's 32 KB instruction cache: with blocks of 8 (43-64 KB of
# straight-line code per shape) the substitution threads spent half their cycles waiting for instructions
# (profiles/r2d_tri_g8_ncu.txt).
TRI_G = int(__import__("os").environ.get("PMC_TRI_G", "4"))      # 4 or 8 (csrc/flow_tri.cu is built for both)
TRI_VERSION = 203
(TRI_D, TRI_H, TRI_L, TRI_T, TRI_NB, TRI_HC, TRI_COL_OUT, TRI_NCOLS, TRI_TSTRIDE, TRI_NCHUNKS, TRI_SLOT_BYTES, TRI_DSLOT_BYTES,
 TRI_TILE_BYTES, TRI_OFF_BLOCKS, TRI_OFF_CHUNKS, TRI_VER, TRI_NSTAGES, TRI_GSIZE, TRI_HEADER) = range(19)
TRI_BLOCK_FIELDS = 10     # k0, nstages, U, W (slots; the K extent of the A tiles is W rounded up to 8), hc (first hidden column), diag offset (floats), diag floats, chunk0, n_urgent, n_chunks
TRI_CHUNK_FIELDS = 8      # a_src (0 = x tile, l = layer-l tile), ks0, nks, N, dcol, first (1: overwrite), offset (floats), flags (1 last urgent, 2 last of block)
TRI_SMEM_BUDGET = 227 * 1024
TRI_BSLOT_TARGET = 24 * 1024
# one update group per block (every later column at once) instead of "next block first, the rest behind it": the MMA
# issuer is a single thread and its instruction stream, not the tensor pipe, bounds the update (profiles/r2k)
TRI_SPLIT_UPDATES = False
TRI_MAX_STAGES = 16       # update-slab ring: as many slots as shared memory allows, at most this (csrc/flow_tri.cu)


def tri_slot(j, s, G, E):
    """column / K index of unit s of in-block group j: four regular columns per group, extras behind them."""
    return 4 * j + s if s < 4 else 4 * G + E * j + (s - 4)


def tri_diag_floats(G: int, U: int) -> int:
    """floats of one block's FFMA weight slab (format walked in lock step by csrc/flow_tri.cu: tri_stage; float4
    granularity).  NR = 4 + E destination units per group; sources come in PAIRS (packed fp32 FMAs: one float2 of
    weights (w[src 2p], w[src 2p+1]) per destination unit):
      stage j:  out bias f4 | out weights of source groups 0..j-1: 2 f4 (+1 f4 extras)
                layer-1 bias | layer-1 weights of x pairs 0..j>>1: NR float2 each
                layers 2, 3: bias | source groups 0..j: NR f4 regular (+ 2 f4 for the extra source)
    plus a tail pad of one group (the kernel prefetches one group ahead)."""
    E = max(U - 4, 0)
    if E > 1:
        raise ValueError("block shapes with more than 5 units per group are not built")
    NR = 4 + E
    nrv = 1 if NR == 4 else 2
    q1 = (2 * NR + 3) // 4
    n = 0
    for j in range(G):
        n += 1 + j * (2 + (1 if E else 0))
        n += nrv + (j // 2 + 1) * q1
        n += 2 * (nrv + (j + 1) * (NR + (2 if E else 0)))
    return 4 * (n + NR + 2)


@dataclass(frozen=True)
class TriLayout:
    tstride: int           # floats per transform in the packed image
    meta: np.ndarray       # int32
    gather: np.ndarray     # int32 [T * tstride]  (pmc_flow_tc_pack codes)
    smem_bytes: int
    blocks: tuple          # per block dict (host-side description, used by the emulator / tests)
    chunks: tuple

    @property
    def numel(self):
        return int(self.gather.size)


def _tri_blocks(D: int, H: int):
    G = TRI_G
    ng = D - 1
    deg = (np.arange(H) % ng) + 1
    hperm = np.argsort(deg, kind="stable")
    gstart = np.searchsorted(deg[hperm], np.arange(1, ng + 2), side="left")
    gsize = np.diff(gstart)
    blocks = []
    hc = 0
    for b in range((D + G - 1) // G):
        k0 = b * G
        nst = min(D, k0 + G) - k0
        groups = [g for g in range(k0 + 1, k0 + nst + 1) if g <= ng]
        U = int(max([gsize[g - 1] for g in groups], default=0))
        E = max(U - 4, 0)
        W = 4 * G + E * G
        unit = np.full(W, -1, np.int64)                  # slot -> original hidden unit
        for g in groups:
            for s in range(int(gsize[g - 1])):
                unit[tri_slot(g - (k0 + 1), s, G, E)] = hperm[gstart[g - 1] + s]
        blocks.append(dict(k0=k0, nst=nst, U=U, E=E, W=W, Kp=(W + 7) // 8 * 8, Nb=(W + 15) // 16 * 16, hc=hc, unit=unit))
        hc += blocks[-1]["Nb"]
    return blocks, hc


def tri_supported(n_dim: int, n_hidden: int, n_layers: int, kind: int) -> bool:
    if kind != KIND_AFFINE or n_layers != 3 or n_dim < 2:
        return False
    blocks, Hc = _tri_blocks(n_dim, n_hidden)
    if any(b["U"] > 5 for b in blocks) or len(blocks) < 2:
        return False
    if n_layers * Hc + (2 * n_dim + 15) // 16 * 16 > 512:
        return False
    try:
        return build_tri(n_dim, n_hidden, n_layers, 1, kind).smem_bytes <= TRI_SMEM_BUDGET
    except ValueError:
        return False


@lru_cache(maxsize=None)
def build_tri(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8) -> TriLayout:
    if kind != KIND_AFFINE or n_layers != 3:
        raise ValueError("the tcgen05 block-triangular sweep is built for affine flows with 3 hidden layers")
    lay = build_layout(n_dim, n_hidden, n_layers, n_transforms, kind, bins)
    D, H, L, T, G = n_dim, n_hidden, n_layers, n_transforms, TRI_G
    blocks, Hc = _tri_blocks(D, H)
    NB = len(blocks)
    if any(b["U"] > 5 for b in blocks) or NB < 2:
        raise ValueError("degree groups too wide (or too few order positions) for the block-triangular sweep")
    col_out = L * Hc
    n_out_cols = (2 * D + 15) // 16 * 16
    ncols = col_out + n_out_cols
    if ncols > 512:
        raise ValueError("accumulators exceed tensor memory (512 columns)")
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in lay.raw_sizes])]).astype(np.int64)
    PLAIN = TC_BIAS_FLAG

    # destination column -> (kind, index) tables shared by every transform
    col_unit = np.full(Hc, -1, np.int64)                 # hidden column -> original unit
    for b in blocks:
        col_unit[b["hc"]:b["hc"] + b["W"]] = b["unit"]

    def transform_parts(t):
        base = t * lay.raw_tstride
        iperm = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
        w = [base + raw_off[2 * l] for l in range(L + 1)]
        bia = [base + raw_off[2 * l + 1] for l in range(L + 1)]
        parts, chunk_rows, block_rows = [], [], []
        off = 0
        for bi, b in enumerate(blocks):
            k0, nst, U, E, W, unit = b["k0"], b["nst"], b["U"], b["E"], b["W"], b["unit"]
            row = 4 + (4 if E else 0)
            # ---- diagonal (FFMA) slab: see tri_diag_floats for the format ----
            NR = 4 + E
            nrv = 1 if NR == 4 else 2
            q1 = (2 * NR + 3) // 4

            def wsrc(widx, width, dst_unit, src_key):
                """raw index of W_widx[dst_unit, src_key] or -1"""
                return -1 if (dst_unit < 0 or src_key < 0) else w[widx] + dst_unit * width + src_key

            d = []
            for j in range(G):
                k = k0 + j
                valid_stage = j < nst
                feat = iperm[k] if valid_stage else -1
                row_s = 2 * feat if valid_stage else -1            # output rows (shift, scale_raw) of this order position
                bo = np.full(4, -1, np.int64)
                if valid_stage:
                    bo[0], bo[1] = bia[L] + row_s, bia[L] + row_s + 1
                d.append(bo)
                for c in range(j):
                    for p in range(2):
                        u0, u1 = unit[4 * c + 2 * p], unit[4 * c + 2 * p + 1]
                        d.append(np.array([wsrc(L, H, row_s, u0), wsrc(L, H, row_s, u1),
                                           wsrc(L, H, row_s + 1 if valid_stage else -1, u0),
                                           wsrc(L, H, row_s + 1 if valid_stage else -1, u1)], np.int64))
                    if E:
                        ue = unit[4 * G + c]
                        d.append(np.array([wsrc(L, H, row_s, ue), wsrc(L, H, row_s + 1 if valid_stage else -1, ue), -1, -1], np.int64))
                own_unit = np.array([unit[tri_slot(j, s_, G, E)] for s_ in range(NR)], np.int64)
                bb = np.full(4 * nrv, -1, np.int64)
                ok = own_unit >= 0
                bb[:NR][ok] = bia[0] + own_unit[ok]
                d.append(bb)
                for q in range(j // 2 + 1):
                    blk = np.full(4 * q1, -1, np.int64)
                    for s_ in range(NR):
                        for h in range(2):
                            i = 2 * q + h
                            if i <= j and k0 + i < D:
                                blk[2 * s_ + h] = wsrc(0, D, own_unit[s_], iperm[k0 + i])
                    d.append(blk)
                for l in range(1, L):
                    bb = np.full(4 * nrv, -1, np.int64)
                    bb[:NR][ok] = bia[l] + own_unit[ok]
                    d.append(bb)
                    for c in range(j + 1):
                        blk = np.full(4 * NR, -1, np.int64)
                        for p in range(2):
                            for s_ in range(NR):
                                for h in range(2):
                                    blk[(p * NR + s_) * 2 + h] = wsrc(l, H, own_unit[s_], unit[4 * c + 2 * p + h])
                        d.append(blk)
                        if E:
                            blk = np.full(8, -1, np.int64)
                            for s_ in range(NR):
                                blk[s_] = wsrc(l, H, own_unit[s_], unit[4 * G + c])
                            d.append(blk)
            d.append(np.full(4 * (NR + 2), -1, np.int64))              # tail pad: the kernel prefetches one group ahead
            d = np.concatenate(d)
            assert len(d) == tri_diag_floats(G, U), (len(d), tri_diag_floats(G, U))
            d = np.where(d >= 0, d | PLAIN, -1)
            diag_off, diag_n = off, len(d)
            parts.append(d)
            off += len(d)
            # ---- update (MMA) slabs: urgent = next block's columns, rest = everything after it ----
            chunk0 = len(chunk_rows)
            n_urgent = 0
            if bi + 1 < NB:
                nxt = blocks[bi + 1]
                # outputs: every remaining output column in ONE urgent update; widths are multiples of 16, so when the
                # remainder is 8 mod 16 the update starts 8 columns early, on this block's own (already consumed) outputs
                rem = n_out_cols - 2 * G * (bi + 1)
                o_start = 2 * G * (bi + 1) - (8 if rem % 16 else 0)
                if TRI_SPLIT_UPDATES:
                    ranges = [("u", nxt["hc"], nxt["hc"] + nxt["Nb"], o_start, n_out_cols)]
                    if bi + 2 < NB:
                        ranges.append(("r", blocks[bi + 2]["hc"], Hc, 0, 0))
                else:
                    ranges = [("u", nxt["hc"], Hc, o_start, n_out_cols)]
                for tag, h0, h1, o0, o1 in ranges:
                    for ld in range(1, L + 2):                          # destination: hidden layer ld, or outputs (L + 1)
                        if ld <= L:
                            c0, c1, dcol = h0, h1, (ld - 1) * Hc + h0
                        else:
                            c0, c1, dcol = o0, o1, col_out + o0
                        N = c1 - c0
                        if N <= 0:
                            continue
                        K = 8 if ld == 1 else b["Kp"]
                        idx = np.full((N, K), -1, np.int64)
                        for n in range(N):
                            if ld <= L:
                                u_dst = col_unit[c0 + n]
                                if u_dst < 0:
                                    continue
                                if ld == 1:
                                    for i in range(min(8, nst)):
                                        idx[n, i] = w[0] + u_dst * D + iperm[k0 + i]
                                else:
                                    ok = np.nonzero(unit >= 0)[0]
                                    idx[n, ok] = w[ld - 1] + u_dst * H + unit[ok]
                            else:
                                col = c0 + n
                                kk, c = col // 2, col % 2
                                if kk >= D or kk < k0 + nst:
                                    continue                      # padding columns / this block's own consumed outputs
                                ok = np.nonzero(unit >= 0)[0]
                                idx[n, ok] = w[L] + (2 * iperm[kk] + c) * H + unit[ok]
                        # split into chunks of whole k-steps that fit a ring slot
                        per_ks = N * 8 * 4 * 2
                        max_ks = max(1, TRI_BSLOT_TARGET // per_ks)
                        ks = 0
                        while ks < K // 8:
                            nks = min(max_ks, K // 8 - ks)
                            sub = idx[:, 8 * ks:8 * (ks + nks)]
                            img = sub.reshape(N, nks * 2, 4).transpose(1, 0, 2).reshape(-1)        # [k/4][N][4]
                            parts.append(img)
                            parts.append(np.where(img >= 0, -(img + 2), -1))
                            chunk_rows.append([0 if ld == 1 else ld - 1, ks, nks, N, dcol, 1 if (bi == 0 and ks == 0) else 0, off, 0])
                            off += 2 * len(img)
                            ks += nks
                    if tag == "u":
                        chunk_rows[-1][7] |= 1
                        n_urgent = len(chunk_rows) - chunk0
                chunk_rows[-1][7] |= 2
            block_rows.append([k0, nst, U, W, b["hc"], diag_off, diag_n, chunk0, n_urgent, len(chunk_rows) - chunk0])
        return np.concatenate(parts), chunk_rows, block_rows

    gathers = []
    for t in range(T):
        gthr, chunk_rows, block_rows = transform_parts(t)
        gathers.append(gthr)
    tstride = len(gathers[0])
    assert all(len(g) == tstride for g in gathers) and tstride % 4 == 0
    chunk_rows = np.asarray(chunk_rows, np.int64).reshape(-1, TRI_CHUNK_FIELDS)
    block_rows = np.asarray(block_rows, np.int64)
    chunk_bytes = chunk_rows[:, 2] * chunk_rows[:, 3] * 64 if len(chunk_rows) else np.zeros(1, np.int64)
    slot_bytes = int((chunk_bytes.max() + 1023) // 1024 * 1024)
    dslot_bytes = int((block_rows[:, 6].max() * 4 + 1023) // 1024 * 1024)
    tile_bytes = int(max(b["Kp"] for b in blocks)) * 512
    if len(chunk_rows) > 96 or NB > 12:
        raise ValueError("too many update chunks / blocks for the kernel's tables")
    fixed = L * 2 * tile_bytes + 2 * 4096 + 2 * dslot_bytes + 8192        # + the kernel's static tables and barriers
    stages = min(TRI_MAX_STAGES, (TRI_SMEM_BUDGET - fixed) // slot_bytes)
    if stages < 2:
        raise ValueError("shared memory budget exceeded")
    smem = fixed + stages * slot_bytes
    meta = np.zeros(TRI_HEADER, np.int64)
    meta[[TRI_D, TRI_H, TRI_L, TRI_T, TRI_NB, TRI_HC, TRI_COL_OUT, TRI_NCOLS, TRI_TSTRIDE, TRI_NCHUNKS, TRI_SLOT_BYTES,
          TRI_DSLOT_BYTES, TRI_TILE_BYTES, TRI_VER, TRI_NSTAGES, TRI_GSIZE]] = \
        [D, H, L, T, NB, Hc, col_out, ncols, tstride, len(chunk_rows), slot_bytes, dslot_bytes, tile_bytes, TRI_VERSION, stages, G]
    meta[TRI_OFF_BLOCKS] = TRI_HEADER
    meta[TRI_OFF_CHUNKS] = TRI_HEADER + block_rows.size
    meta = np.concatenate([meta, block_rows.reshape(-1), chunk_rows.reshape(-1)])
    gather = np.concatenate(gathers)
    assert np.abs(gather).max() < 2 ** 31 and meta.max() < 2 ** 31
    return TriLayout(tstride, meta.astype(np.int32), gather.astype(np.int32), int(smem),
                     tuple({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in b.items()} for b in blocks),
                     tuple(map(tuple, chunk_rows.tolist())))
