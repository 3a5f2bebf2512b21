"""Host-side layout of the tcgen05 block-triangular sweep (csrc/flow_tri.cu): Flow.inverse / Flow.forward of affine
(zuko MAF) and rational-quadratic-spline (zuko NSF, the reference's default presets) flows of ANY width.

The degree-ordered sweep (made_layout.py) is a nonlinear forward substitution.  Order positions are cut into BLOCKS of
TRI_G = 4; the hidden units born in a block (its degree groups) form one K-slab.  Blocks are grouped into WINDOWS: the
running pre-activations ("accumulators") of a window's units -- three hidden layers and the outputs -- live in tensor
memory (512 fp32 columns per 128-particle tile), so a window holds as many blocks as fit.

  * inside a window the schedule is right-looking: when a block is finished, its activations (A tiles [128 x K], TF32
    hi / lo images in shared memory) update the accumulators of all later columns of the window with one tcgen05.mma
    group per layer;
  * when a window is finished and another follows, the next window's accumulators are INITIALISED left-looking from the
    activations of every earlier block, which the kernel also keeps in a per-CTA scratch area in global memory (plain
    fp32 in the K-major chunk layout of the shared-memory tiles, so a K range is one contiguous bulk copy): a stream of
    [A chunk from scratch | B chunk of weights] pairs through the same shared-memory ring, the A chunk split into its
    TF32 hi / lo images in place by the particle threads;
  * the dependencies inside a block run as fp32 FMAs, one thread per particle (the in-block slab).

A flow whose accumulators fit one window (affine, D <= 36 at the preset widths) never touches the scratch area.

Outputs per order position: P = 2 tensor-memory columns for affine heads (shift, raw log-scale), P = 24 for spline heads (the
23 = 3 * 8 - 1 spline parameters of zuko's NSF padded by one zero column); a window holds at most 256 output columns
(the widest tcgen05.mma), i.e. two blocks of a spline flow.

Packed image per transform (floats; gather codes of pmc_flow_tc_pack: >= 0 hi, -(g+2) lo, g | 2^30 plain, -1 zero):
    [in-block slab of block 0] ... [in-block slab of block NB-1]
    [chunk stream in consumption order: for every block that is not the last of its window the four update slabs
     (layer 1, 2, 3, outputs); for the last block of a window that is not the last window the init slabs of the next
     window: for layer 1, 2, 3, outputs: K chunks of TRI_KC]
Every update / init slab is B = [N dest columns x K source slots] in the no-swizzle K-major UMMA layout [K/4][N][4],
hi image then lo image.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache

import numpy as np

from .made_layout import KIND_AFFINE, KIND_RQS, TC_BIAS_FLAG, build_layout

TRI_G = 4
TRI_KC = 16               # K extent of one init chunk (A from the scratch area + B weights); 8 when 16 leaves < 4 ring slots
TRI_VERSION = 302
(TRI_VER, TRI_D, TRI_H, TRI_L, TRI_T, TRI_GSIZE, TRI_NB, TRI_NW, TRI_TSTRIDE, TRI_KCHUNK, TRI_SLOT_BYTES, TRI_DSLOT_BYTES,
 TRI_TILE_BYTES, TRI_NSTAGES, TRI_KH_TOTAL, TRI_KX_TOTAL, TRI_WS_FLOATS, TRI_OFF_BLOCKS, TRI_OFF_WINDOWS, TRI_SMEM_BYTES,
 TRI_NCOLS, TRI_CHUNK_OFF, TRI_KIND, TRI_KCHUNK_OUT, TRI_HEADER) = range(25)
# per block
(TB_K0, TB_NST, TB_NR, TB_W, TB_KP, TB_WIN, TB_WC, TB_OC, TB_DOFF, TB_DN, TB_KS, TB_UPD_N, TB_UPD_DCOL, TB_OUT_N, TB_OUT_DCOL,
 TB_FLAGS, TB_FIELDS) = range(17)
# per window
(TW_B0, TW_NB, TW_WP, TW_OP, TW_COL_OUT, TW_KH, TW_KX, TW_PAD, TW_FIELDS) = range(9)
TRI_SMEM_BUDGET = 227 * 1024
TRI_STATIC_SMEM = 14 * 1024       # barriers, tables and per-block issue records of the kernel (sizeof(TriShared))
TRI_MAX_STAGES = 8
TRI_MAX_BLOCKS = 64
TRI_MAX_WINDOWS = 32
TRI_P_RQS = 24             # output columns per order position of a spline flow (23 parameters + one zero column)


def tri_slot(j, s, G, E):
    """column / K index of unit s of in-block group j: four regular columns per group, the E extras behind them."""
    return 4 * j + s if s < 4 else 4 * G + E * j + (s - 4)


def _shape(NR):
    E = NR - 4
    return dict(E=E, nrv=1 if NR == 4 else 2, q1=(2 * NR + 3) // 4, xv=(0, 2, 3)[E], ogx=(0, 1, 1)[E])


def tri_out_cols(kind: int) -> int:
    """tensor-memory columns per order position for the univariate head's parameters"""
    return 2 if kind == KIND_AFFINE else TRI_P_RQS


def tri_diag_floats(G: int, NR: int, kind: int = KIND_AFFINE) -> int:
    """floats of one block's FFMA weight slab (walked in lock step by csrc/flow_tri.cu: tri_stage; float4 granularity).
    NR = 4 + E destination units per group; sources come in PAIRS (packed fp32 FMAs: one float2 of weights
    (w[src 2p], w[src 2p+1]) per destination unit):
      stage j:  out bias f4 | out weights of source groups 0..j-1: 2 f4 (+1 f4 for the extra unit(s))
                [spline heads: out bias 6 f4 (24 parameters) | per source group and source unit (4 regular, then the
                 extras) 6 f4 = that unit's weight into each of the 24 parameters]
                layer-1 bias | layer-1 weights of x pairs 0..j>>1: NR float2 each
                layers 2, 3: bias | source groups 0..j: NR f4 regular + the extras (E = 1: NR floats in 2 f4;
                                                                                  E = 2: NR float2 in 3 f4)."""
    sh = _shape(NR)
    n = 0
    for j in range(G):
        n += (1 + j * (2 + sh["ogx"])) if kind == KIND_AFFINE else (6 + j * 6 * (4 + sh["E"]))
        n += sh["nrv"] + (j // 2 + 1) * sh["q1"]
        n += 2 * (sh["nrv"] + (j + 1) * (NR + sh["xv"]))
    return 4 * n


@dataclass(frozen=True)
class TriLayout:
    tstride: int           # floats per transform in the packed image
    meta: np.ndarray       # int32: header, block table, window table
    gather: np.ndarray     # int32 [T * tstride]  (pmc_flow_tc_pack codes)
    smem_bytes: int
    ws_floats: int         # scratch floats per CTA (0: the flow fits one window)
    blocks: tuple
    windows: tuple

    @property
    def numel(self):
        return int(self.gather.size)


def _r(n, m):
    return (int(n) + m - 1) // m * m


def _tri_blocks(D: int, H: int, P: int = 2):
    """blocks (with their unit maps) and windows of a (D, H) masked MLP with P output columns per order position"""
    G = TRI_G
    PG = P * G
    ng = D - 1
    deg = (np.arange(H) % ng) + 1
    hperm = np.argsort(deg, kind="stable")
    gstart = np.searchsorted(deg[hperm], np.arange(1, ng + 2), side="left")
    gsize = np.diff(gstart)
    blocks = []
    ks = 0
    for b in range((D + G - 1) // G):
        k0 = b * G
        nst = min(D, k0 + G) - k0
        groups = [g for g in range(k0 + 1, k0 + nst + 1) if g <= ng]
        U = int(max([gsize[g - 1] for g in groups], default=0))
        NR = max(U, 4)
        E = NR - 4
        W = 4 * G + E * G
        unit = np.full(W, -1, np.int64)                  # slot -> original hidden unit
        for g in groups:
            for s in range(int(gsize[g - 1])):
                unit[tri_slot(g - (k0 + 1), s, G, E)] = hperm[gstart[g - 1] + s]
        blocks.append(dict(k0=k0, nst=nst, U=U, NR=NR, E=E, W=W, Kp=_r(W, 8), ks=ks, unit=unit))
        ks += blocks[-1]["Kp"]
    # windows: greedy packing of blocks into 512 tensor-memory columns (3 hidden layers + outputs)
    windows, b0 = [], 0
    while b0 < len(blocks):
        nb, wsum = 0, 0
        while b0 + nb < len(blocks):
            w2 = wsum + blocks[b0 + nb]["W"]
            if 3 * _r(w2, 16) + _r(PG * (nb + 1), 16) > 512 or _r(PG * (nb + 1), 16) > 256:
                break
            wsum, nb = w2, nb + 1
        if nb == 0:
            raise ValueError("a single block exceeds tensor memory")
        windows.append(dict(b0=b0, nb=nb, Wp=_r(wsum, 16), Op=_r(PG * nb, 16), Kh=blocks[b0]["ks"], Kx=8 * b0))
        wc = 0
        for i in range(nb):
            blk = blocks[b0 + i]
            blk.update(win=len(windows) - 1, wc=wc, oc=PG * i, last_in_win=(i == nb - 1))
            wc += blk["W"]
        b0 += nb
    for w in windows:
        w["col_out"] = 3 * w["Wp"]
    for blk in blocks:
        w = windows[blk["win"]]
        if blk["last_in_win"]:
            blk.update(upd_N=0, upd_dcol=0, out_N=0, out_dcol=0)
        else:
            nxt = blk["wc"] + blk["W"]                   # first later hidden column of the window
            start = nxt // 16 * 16                       # MMA widths are multiples of 16 (Wp is one): start early, on consumed columns
            onxt = blk["oc"] + PG
            ostart = onxt // 16 * 16
            blk.update(upd_N=w["Wp"] - start, upd_dcol=start, out_N=w["Op"] - ostart, out_dcol=ostart)
    return blocks, windows, ks


def tri_supported(n_dim: int, n_hidden: int, n_layers: int, kind: int) -> bool:
    if kind not in (KIND_AFFINE, KIND_RQS) or n_layers != 3 or n_dim < 2:
        return False
    try:
        build_tri(n_dim, n_hidden, n_layers, 1, kind)
        return True
    except ValueError:
        return False


@lru_cache(maxsize=None)
def build_tri(n_dim: int, n_hidden: int, n_layers: int, n_transforms: int, kind: int, bins: int = 8) -> TriLayout:
    if kind not in (KIND_AFFINE, KIND_RQS) or n_layers != 3 or (kind == KIND_RQS and bins != 8):
        raise ValueError("the tcgen05 block-triangular sweep is built for affine / 8-bin spline flows with 3 hidden layers")
    lay = build_layout(n_dim, n_hidden, n_layers, n_transforms, kind, bins)
    D, H, L, T, G = n_dim, n_hidden, n_layers, n_transforms, TRI_G
    P, total = tri_out_cols(kind), lay.total
    PG = P * G
    blocks, windows, kh_total = _tri_blocks(D, H, P)
    NB, NW = len(blocks), len(windows)
    if any(b["U"] > 6 for b in blocks) or NB < 2:
        raise ValueError("degree groups too wide (or too few order positions) for the block-triangular sweep")
    if NB > TRI_MAX_BLOCKS or NW > TRI_MAX_WINDOWS:
        raise ValueError("too many blocks / windows for the kernel's tables")
    raw_off = np.concatenate([[0], np.cumsum([int(np.prod(s)) for s in lay.raw_sizes])]).astype(np.int64)
    PLAIN = TC_BIAS_FLAG
    # ---- sizes of the shared-memory ring (decide the K extent of an init chunk before the chunk stream is laid out) ----
    max_upd = max([max(8, b["Kp"]) * max(b["upd_N"], b["out_N"]) * 8 for b in blocks if not b["last_in_win"]], default=0)
    dslot_bytes = _r(max(tri_diag_floats(G, b["NR"], kind) for b in blocks) * 4, 1024)
    tile_bytes = max(b["Kp"] for b in blocks) * 512
    fixed = L * 2 * tile_bytes + 2 * 4096 + 2 * dslot_bytes + TRI_STATIC_SMEM
    # K extent of an init chunk for the hidden layers (KC) and for the outputs (KCO): a chunk is [A: 128 rows x K fp32 | B: N x K
    # hi + lo]; spline flows have wide output slabs (N = 192), so their output chunks are shorter than their hidden ones.
    # Every chunk costs one hand-off between producer, splitting threads and issuer: take the longest that leaves >= 4 slots.
    for KC, KCO in (((TRI_KC, TRI_KC),) if kind == KIND_AFFINE else ((32, 16), (16, 16), (16, 8), (8, 8))):
        max_init = max([max(KC * (1024 + w["Wp"] * 8), KCO * (1024 + w["Op"] * 8)) for w in windows[1:]], default=0)
        slot_bytes = _r(max(max_upd, max_init, 1024), 1024)
        stages = min(TRI_MAX_STAGES, (TRI_SMEM_BUDGET - fixed) // slot_bytes)
        if stages >= 4:
            break
    if stages < 2:
        raise ValueError("shared memory budget exceeded")
    # global slot axes: hidden slot -> unit (all blocks), x slot -> order position
    hslot_unit = np.full(kh_total, -1, np.int64)
    for b in blocks:
        hslot_unit[b["ks"]:b["ks"] + b["W"]] = b["unit"]
    kx_total = 8 * NB
    xslot_order = np.full(kx_total, -1, np.int64)
    for bi, b in enumerate(blocks):
        xslot_order[8 * bi:8 * bi + b["nst"]] = b["k0"] + np.arange(b["nst"])

    def k_major(idx):
        """[N, K] index matrix -> hi image, lo image in [K/4][N][4] order"""
        N, K = idx.shape
        img = idx.reshape(N, K // 4, 4).transpose(1, 0, 2).reshape(-1)
        return [img, np.where(img >= 0, -(img + 2), -1)]

    def transform_parts(t):
        base = t * lay.raw_tstride
        iperm = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
        w = [base + raw_off[2 * l] for l in range(L + 1)]
        bia = [base + raw_off[2 * l + 1] for l in range(L + 1)]

        def wsrc(widx, width, dst, src):
            return -1 if (dst < 0 or src < 0) else w[widx] + dst * width + src

        parts, off = [], 0
        diag_tab = []
        # ---- in-block (FFMA) slabs ----
        for b in blocks:
            k0, nst, NR, E, unit = b["k0"], b["nst"], b["NR"], b["E"], b["unit"]
            sh = _shape(NR)
            d = []
            for j in range(G):
                valid = j < nst
                row_s = 2 * iperm[k0 + j] if valid else -1
                row_r = row_s + 1 if valid else -1
                if kind == KIND_RQS:
                    rows = np.full(P, -1, np.int64)
                    if valid:
                        rows[:total] = total * iperm[k0 + j] + np.arange(total)
                    d.append(np.where(rows >= 0, bia[L] + rows, -1))
                    for c in range(j):
                        for src in [unit[4 * c + u_] for u_ in range(4)] + [unit[4 * G + E * c + e_] for e_ in range(E)]:
                            d.append(np.array([wsrc(L, H, r_, src) for r_ in rows], np.int64))
                else:
                    bo = np.full(4, -1, np.int64)
                    if valid:
                        bo[0], bo[1] = bia[L] + row_s, bia[L] + row_r
                    d.append(bo)
                for c in range(j if kind == KIND_AFFINE else 0):
                    for p in range(2):
                        u0, u1 = unit[4 * c + 2 * p], unit[4 * c + 2 * p + 1]
                        d.append(np.array([wsrc(L, H, row_s, u0), wsrc(L, H, row_s, u1), wsrc(L, H, row_r, u0), wsrc(L, H, row_r, u1)], np.int64))
                    if E == 1:
                        ue = unit[4 * G + c]
                        d.append(np.array([wsrc(L, H, row_s, ue), wsrc(L, H, row_r, ue), -1, -1], np.int64))
                    elif E == 2:
                        e0, e1 = unit[4 * G + 2 * c], unit[4 * G + 2 * c + 1]
                        d.append(np.array([wsrc(L, H, row_s, e0), wsrc(L, H, row_s, e1), wsrc(L, H, row_r, e0), wsrc(L, H, row_r, e1)], np.int64))
                own = np.array([unit[tri_slot(j, s_, G, E)] for s_ in range(NR)], np.int64)
                ok = own >= 0
                bb = np.full(4 * sh["nrv"], -1, np.int64)
                bb[:NR][ok] = bia[0] + own[ok]
                d.append(bb)
                for q in range(j // 2 + 1):
                    blk = np.full(4 * sh["q1"], -1, np.int64)
                    for s_ in range(NR):
                        for h in range(2):
                            i = 2 * q + h
                            if i <= j and k0 + i < D:
                                blk[2 * s_ + h] = wsrc(0, D, own[s_], iperm[k0 + i])
                    d.append(blk)
                for l in range(1, L):
                    bb = np.full(4 * sh["nrv"], -1, np.int64)
                    bb[:NR][ok] = bia[l] + own[ok]
                    d.append(bb)
                    for c in range(j + 1):
                        blk = np.full(4 * NR, -1, np.int64)
                        for p in range(2):
                            for s_ in range(NR):
                                for h in range(2):
                                    blk[(p * NR + s_) * 2 + h] = wsrc(l, H, own[s_], unit[4 * c + 2 * p + h])
                        d.append(blk)
                        if E == 1:
                            blk = np.full(8, -1, np.int64)
                            for s_ in range(NR):
                                blk[s_] = wsrc(l, H, own[s_], unit[4 * G + c])
                            d.append(blk)
                        elif E == 2:
                            blk = np.full(12, -1, np.int64)
                            for s_ in range(NR):
                                for h in range(2):
                                    blk[2 * s_ + h] = wsrc(l, H, own[s_], unit[4 * G + 2 * c + h])
                            d.append(blk)
            d = np.concatenate(d)
            assert len(d) == tri_diag_floats(G, NR, kind), (len(d), tri_diag_floats(G, NR, kind))
            diag_tab.append((off, len(d)))
            parts.append(np.where(d >= 0, d | PLAIN, -1))
            off += len(d)
        chunk_off = off

        # destination columns of a window: hidden column -> unit, output column -> (order position, which)
        def win_maps(wd):
            col_unit = np.full(wd["Wp"], -1, np.int64)
            col_k0 = np.full(wd["Wp"], 10 ** 9, np.int64)            # first order position of the column's block
            out_row = np.full(wd["Op"], -1, np.int64)
            for i in range(wd["nb"]):
                bb = blocks[wd["b0"] + i]
                col_unit[bb["wc"]:bb["wc"] + bb["W"]] = bb["unit"]
                col_k0[bb["wc"]:bb["wc"] + bb["W"]] = bb["k0"]
                for j in range(bb["nst"]):
                    out_row[bb["oc"] + P * j:bb["oc"] + P * j + total] = total * iperm[bb["k0"] + j] + np.arange(total)
            return col_unit, col_k0, out_row

        def b_matrix(op, dst_cols, src_units=None, src_orders=None, col_unit=None, out_row=None):
            """[N, K] raw indices: op 1 = layer 1 (sources: order positions), 2 / 3 = hidden layer, 4 = outputs"""
            N = len(dst_cols)
            K = len(src_orders) if op == 1 else len(src_units)
            idx = np.full((N, K), -1, np.int64)
            for n_, c in enumerate(dst_cols):
                if c < 0:
                    continue
                if op == 1:
                    u = col_unit[c]
                    if u >= 0:
                        ok = src_orders >= 0
                        idx[n_, ok] = w[0] + u * D + iperm[src_orders[ok]]
                elif op in (2, 3):
                    u = col_unit[c]
                    if u >= 0:
                        ok = src_units >= 0
                        idx[n_, ok] = w[op - 1] + u * H + src_units[ok]
                else:
                    r = out_row[c]
                    if r >= 0:
                        ok = src_units >= 0
                        idx[n_, ok] = w[L] + r * H + src_units[ok]
            return idx

        # ---- chunk stream ----
        for bi, b in enumerate(blocks):
            wd = windows[b["win"]]
            col_unit, col_k0, out_row = win_maps(wd)
            if not b["last_in_win"]:
                # update slabs: destination = later columns of the window (columns before them, reached only because
                # widths are multiples of 16, get zero weights: they belong to consumed blocks)
                nxt, onxt = b["wc"] + b["W"], b["oc"] + PG
                hcols = np.array([c if c >= nxt else -1 for c in range(b["upd_dcol"], wd["Wp"])], np.int64)
                ocols = np.array([c if c >= onxt else -1 for c in range(b["out_dcol"], wd["Op"])], np.int64)
                xo = np.full(8, -1, np.int64)
                xo[:b["nst"]] = b["k0"] + np.arange(b["nst"])
                upad = np.full(b["Kp"], -1, np.int64)
                upad[:b["W"]] = b["unit"]
                for op in (1, 2, 3, 4):
                    idx = b_matrix(op, hcols if op < 4 else ocols, src_units=upad, src_orders=xo, col_unit=col_unit, out_row=out_row)
                    for img in k_major(idx):
                        parts.append(img)
                        off += len(img)
            elif bi + 1 < NB:
                # init slabs of the next window: every earlier slot -> all of its columns, K chunks of KC
                nw = windows[b["win"] + 1]
                ncol_unit, _, nout_row = win_maps(nw)
                hcols = np.arange(nw["Wp"])
                ocols = np.arange(nw["Op"])
                for op in (1, 2, 3, 4):
                    ktot = nw["Kx"] if op == 1 else nw["Kh"]
                    idx = b_matrix(op, hcols if op < 4 else ocols, src_units=hslot_unit[:ktot], src_orders=xslot_order[:ktot],
                                   col_unit=ncol_unit, out_row=nout_row)
                    kc_ = KC if op < 4 else KCO
                    for k_ in range(0, ktot, kc_):
                        for img in k_major(idx[:, k_:min(k_ + kc_, ktot)]):
                            parts.append(img)
                            off += len(img)
        return np.concatenate(parts), diag_tab, chunk_off

    gathers = []
    for t in range(T):
        gthr, diag_tab, chunk_off = transform_parts(t)
        gathers.append(gthr)
    tstride = len(gathers[0])
    assert all(len(g) == tstride for g in gathers) and tstride % 4 == 0 and chunk_off % 4 == 0

    # ---- sizes ----
    assert dslot_bytes >= max(dn for _, dn in diag_tab) * 4
    smem = fixed + stages * slot_bytes
    ncols = max(3 * w["Wp"] + w["Op"] for w in windows)
    assert ncols <= 512
    ws_floats = 0 if NW == 1 else (kx_total + 3 * kh_total) * 128          # plain fp32; split into hi / lo on the way back in

    block_rows = np.zeros((NB, TB_FIELDS), np.int64)
    for bi, b in enumerate(blocks):
        block_rows[bi] = [b["k0"], b["nst"], b["NR"], b["W"], b["Kp"], b["win"], b["wc"], b["oc"], diag_tab[bi][0], diag_tab[bi][1], b["ks"],
                          b["upd_N"], b["upd_dcol"], b["out_N"], b["out_dcol"], (1 if b["last_in_win"] else 0) | (2 if bi == NB - 1 else 0)]
    win_rows = np.zeros((NW, TW_FIELDS), np.int64)
    for wi, w in enumerate(windows):
        win_rows[wi] = [w["b0"], w["nb"], w["Wp"], w["Op"], w["col_out"], w["Kh"], w["Kx"], 0]
    meta = np.zeros(TRI_HEADER, np.int64)
    meta[[TRI_VER, TRI_D, TRI_H, TRI_L, TRI_T, TRI_GSIZE, TRI_NB, TRI_NW, TRI_TSTRIDE, TRI_KCHUNK, TRI_SLOT_BYTES, TRI_DSLOT_BYTES,
          TRI_TILE_BYTES, TRI_NSTAGES, TRI_KH_TOTAL, TRI_KX_TOTAL, TRI_WS_FLOATS, TRI_SMEM_BYTES, TRI_NCOLS, TRI_CHUNK_OFF, TRI_KIND, TRI_KCHUNK_OUT]] = \
        [TRI_VERSION, D, H, L, T, G, NB, NW, tstride, KC, slot_bytes, dslot_bytes, tile_bytes, stages, kh_total, kx_total, ws_floats,
         smem, ncols, chunk_off, kind, KCO]
    meta[TRI_OFF_BLOCKS] = TRI_HEADER
    meta[TRI_OFF_WINDOWS] = TRI_HEADER + block_rows.size
    meta = np.concatenate([meta, block_rows.reshape(-1), win_rows.reshape(-1)])
    gather = np.concatenate(gathers)
    assert np.abs(gather).max() < 2 ** 31 and meta.max() < 2 ** 31
    return TriLayout(tstride, meta.astype(np.int32), gather.astype(np.int32), int(smem), int(ws_floats),
                     tuple({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in b.items()} for b in blocks),
                     tuple(dict(w) for w in windows))
