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
    MADE mask folded in (masked entries are 0).  Behind the chunks of a transform, at ``bias_off``, the plain fp32
    biases ``[L][H]`` + ``[Nout]`` (added by the epilogue threads out of shared memory).  ``gather`` codes:
    >= 0 hi(raw[g]); -(g+2) lo(raw[g]); g | 2^30 plain copy; -1 zero."""
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
                max_chunk = max(max_chunk, 2 * kc * N * 4)
                if t == 0:
                    n_chunks += 1
        w_floats = int(sum(len(a) for a in parts))
        for l in range(L + 1):                                       # plain fp32 biases behind the chunks: [L][H] then [Nout]
            N_true, N = (D * total, Nout) if l == L else (H, H)
            bb = np.full(N, -1, np.int64)
            bb[:N_true] = (base_r + raw_off[2 * l + 1] + np.arange(N_true)) | TC_BIAS_FLAG
            parts.append(bb)
        parts_all.append(np.concatenate(parts))
    tstride = len(parts_all[0])
    assert tstride % 4 == 0 and w_floats % 4 == 0 and tstride == w_floats + L * H + Nout
    gather = np.concatenate(parts_all)
    slot_bytes = (max_chunk + 1023) // 1024 * 1024
    meta = np.zeros(TC_LEN, np.int64)
    meta[[TC_D, TC_H, TC_L, TC_T, TC_KIND, TC_KX, TC_NOUT, TC_TSTRIDE, TC_BIAS_OFF, TC_NCHUNKS, TC_SLOT_BYTES, TC_VERSION]] = \
        [D, H, L, T, kind, Kx, Nout, tstride, w_floats, n_chunks, slot_bytes, 101]
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
