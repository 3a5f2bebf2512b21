"""Numpy emulation of csrc/flow_sweep.cu's degree-ordered sweep over the packed MADE layout.
Test helper only: lets the CPU suite check pocomc_b200.made_layout (gather map + meta tables)
and the sweep algorithm itself against the zuko oracle without a GPU."""
import math

import numpy as np

from pocomc_b200 import made_layout as ML


def pack(layout, raw):
    raw = np.asarray(raw, np.float32)
    g = layout.gather
    return np.where(g >= 0, raw[np.maximum(g, 0)], np.float32(0)).astype(np.float32)


def _softclip(a, ls):
    return a / (1 + np.abs(a / ls))


def affine(phi, v, inverse):
    ls = _softclip(phi[:, 1], np.float32(math.log(1e-3)))
    if inverse:
        return (v - phi[:, 0]) / np.exp(ls), ls
    return v * np.exp(ls) + phi[:, 0], ls


def rqs(phi, v, inverse, bins=8, bound=5.0):
    n = len(v)
    L = np.float32(math.log(1e-3))
    w = _softclip(phi[:, :bins], L / 2)           # w / (1 + |2w / log slope|)
    h = _softclip(phi[:, bins:2 * bins], L / 2)
    d = _softclip(phi[:, 2 * bins:3 * bins - 1], L)

    def knots(a):
        e = np.exp(a - a.max(1, keepdims=True))
        sm = (e / e.sum(1, keepdims=True)).astype(np.float32)
        c = np.concatenate([np.zeros((n, 1)), np.cumsum(sm.astype(np.float64), 1)], 1).astype(np.float32)
        return (np.float32(bound) * (2 * c - 1)).astype(np.float32)

    hx, hy = knots(w), knots(h)
    dv = np.exp(np.concatenate([np.zeros((n, 1), np.float32), d, np.zeros((n, 1), np.float32)], 1)).astype(np.float32)
    seq = hy if inverse else hx
    k = (seq < v[:, None]).sum(1) - 1
    mask = (k >= 0) & (k < bins)
    k = k % bins
    r = np.arange(n)
    x0, x1, y0, y1, d0, d1 = hx[r, k], hx[r, k + 1], hy[r, k], hy[r, k + 1], dv[r, k], dv[r, k + 1]
    s = (y1 - y0) / (x1 - x0)
    if inverse:
        y_ = mask * (v - y0)
        a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2 * s)
        b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2 * s)
        c = -s * y_
        z = 2 * c / (-b - np.sqrt(b * b - 4 * a * c))
        x = np.where(mask, x0 + z * (x1 - x0), v)
        out = x
    else:
        x = v
    z = mask * (x - x0) / (x1 - x0)
    jac = s * s * (2 * s * z * (1 - z) + d0 * (1 - z) ** 2 + d1 * z * z) / (s + (d0 + d1 - 2 * s) * z * (1 - z)) ** 2
    ladj = np.log(jac) * mask
    if not inverse:
        out = np.where(mask, y0 + (y1 - y0) * (s * z * z + d0 * z * (1 - z)) / (s + (d0 + d1 - 2 * s) * z * (1 - z)), v)
    return out.astype(np.float32), ladj.astype(np.float32)


def sweep(layout, packed, v, inverse):
    """v [N, D] f32 -> (out [N, D] f32, ladj [N] f32).  forward: data->latent, ladj = log|dz/dx|;
    inverse: latent->data, ladj = log|dx/dz| (= -forward ladj at the solution)."""
    m = layout.meta.astype(np.int64)
    D, H, L, T, kind, total, tp = (int(m[i]) for i in (ML.M_D, ML.M_H, ML.M_L, ML.M_T, ML.M_KIND, ML.M_TOTAL, ML.M_TP))
    ng, tstride = int(m[ML.M_NG]), int(m[ML.M_TSTRIDE])
    gstart = m[m[ML.M_OFF_GSTART]:][:ng + 1]
    nchunk = m[m[ML.M_OFF_NCHUNK]:][:ng]
    slot = m[m[ML.M_OFF_SLOT]:][:ng + 1]
    off_w0 = m[m[ML.M_OFF_W0]:][:ng]
    off_wh = m[m[ML.M_OFF_WH]:][:max(L - 1, 1) * ng].reshape(-1, ng)
    off_wo = m[m[ML.M_OFF_WO]:][:D]
    off_bh = m[m[ML.M_OFF_BH]:][:max(L - 1, 1)]
    off_b0, off_bo = int(m[ML.M_OFF_B0]), int(m[ML.M_OFF_BO])
    cur = np.array(v, np.float32, copy=True)
    n = len(cur)
    ladj = np.zeros(n, np.float32)
    uni = affine if kind == ML.KIND_AFFINE else rqs
    for t in (range(T - 1, -1, -1) if inverse else range(T)):
        P = packed[t * tstride:(t + 1) * tstride]
        rev = t % 2 == 1
        xs = np.zeros((D, n), np.float32)                # data-side values by order position
        act = np.zeros((L, H, n), np.float32)
        out = np.empty_like(cur)
        for k in range(D):
            feat = D - 1 - k if rev else k
            Ek = int(gstart[k]) if k >= 1 else 0          # units with degree <= k  (gstart[k] = start of group k+1)
            W = P[off_wo[k]: off_wo[k] + Ek * tp].reshape(Ek, tp)
            phi = (act[L - 1, :Ek].T @ W + P[off_bo + k * tp: off_bo + (k + 1) * tp]).astype(np.float32)
            res, l = uni(phi[:, :total], cur[:, feat], inverse)
            xs[k] = res if inverse else cur[:, feat]
            out[:, feat] = res
            ladj = (ladj - l if inverse else ladj + l).astype(np.float32)
            g = k + 1
            if g > ng or gstart[g] == gstart[g - 1]:
                continue
            gs, ge, wd = int(gstart[g - 1]), int(gstart[g]), 4 * int(nchunk[g - 1])
            nu = ge - gs
            W = P[off_w0[g - 1]: off_w0[g - 1] + g * wd].reshape(g, wd)
            b = P[off_b0 + slot[g - 1]: off_b0 + slot[g - 1] + wd]
            h = np.maximum((xs[:g].T @ W + b).astype(np.float32), 0)[:, :nu]
            act[0, gs:ge] = h.T
            for l_ in range(1, L):
                W = P[off_wh[l_ - 1, g - 1]: off_wh[l_ - 1, g - 1] + ge * wd].reshape(ge, wd)
                b = P[off_bh[l_ - 1] + slot[g - 1]: off_bh[l_ - 1] + slot[g - 1] + wd]
                r = (act[l_ - 1, :ge].T @ W + b).astype(np.float32)[:, :nu] + act[l_ - 1, gs:ge].T
                act[l_, gs:ge] = np.maximum(r, 0).T
        cur = out
    return cur, ladj


def pack_stream(stream, raw):
    raw = np.asarray(raw, np.float32)
    g = stream.gather
    return np.where(g >= 0, raw[np.maximum(g, 0)], np.float32(0)).astype(np.float32)


def sweep_stream(layout, stream, packed, x, inverse):
    """Emulates made_sweep_stream_kernel: walks the consumption-order stream with a running offset."""
    m = stream.meta
    D, H, L, T, ng, tp = (int(m[i]) for i in (ML.M_D, ML.M_H, ML.M_L, ML.M_T, ML.M_NG, ML.M_TP))
    gstart = m[m[ML.M_OFF_GSTART]:m[ML.M_OFF_GSTART] + ng + 1]
    nchunk = m[m[ML.M_OFF_NCHUNK]:m[ML.M_OFF_NCHUNK] + ng]
    chunks = m[m[ML.M_OFF_CHUNKS]:m[ML.M_OFF_CHUNKS] + 4 * m[ML.M_NCHUNKS]].reshape(-1, 4)
    uni = affine if layout.kind == ML.KIND_AFFINE else rqs
    cur = np.array(x, np.float32, copy=True)
    n = len(cur)
    ladj = np.zeros(n, np.float32)
    for tt in range(T):
        t = T - 1 - tt if inverse else tt
        xs = np.zeros((n, D), np.float32)
        act = np.zeros((L, n, H), np.float32)
        for k0, k1, off, cnt in chunks:
            w = packed[t * stream.tstride + off: t * stream.tstride + off + cnt]
            pos = 0
            for k in range(k0, k1):
                feat = D - 1 - k if t % 2 else k
                ek = int(gstart[k])
                phi = np.zeros((n, tp), np.float32)
                p16 = lambda v: (v + 15) // 16 * 16
                for c in range(tp // 4):
                    slab = w[pos:pos + 4 * p16(ek)].reshape(p16(ek), 4); pos += 4 * p16(ek)
                    assert not slab[ek:].any()
                    phi[:, 4 * c:4 * c + 4] = act[L - 1][:, :ek] @ slab[:ek]
                phi += w[pos:pos + tp]; pos += tp
                v = cur[:, feat].copy()
                res, l = uni(phi, v, inverse)
                ladj = ladj - l if inverse else ladj + l
                xs[:, k] = res if inverse else v
                cur[:, feat] = res
                g = k + 1
                if g > ng or gstart[g] == gstart[g - 1]:
                    continue
                gs, ge, nch = int(gstart[g - 1]), int(gstart[g]), int(nchunk[g - 1])
                for l_ in range(L):
                    nrows = g if l_ == 0 else ge
                    src = xs[:, :g] if l_ == 0 else act[l_ - 1][:, :ge]
                    pre = np.zeros((n, 4 * nch), np.float32)
                    for c in range(nch):
                        slab = w[pos:pos + 4 * p16(nrows)].reshape(p16(nrows), 4); pos += 4 * p16(nrows)
                        assert not slab[nrows:].any()
                        pre[:, 4 * c:4 * c + 4] = src @ slab[:nrows]
                    pre += w[pos:pos + 4 * nch]; pos += 4 * nch
                    pre = pre[:, :ge - gs]
                    if l_ > 0:
                        pre = pre + act[l_ - 1][:, gs:ge]
                    act[l_][:, gs:ge] = np.maximum(pre, 0)
            assert pos == cnt
    return cur, ladj


# ---------------------------------------------------------------------------------------------
# block-triangular sweep (csrc/flow_tri.cu): numpy emulation walking the SAME packed image and tables
# ---------------------------------------------------------------------------------------------
def pack_tri(tri, raw):
    """pmc_flow_tc_pack on the host: >= 0 hi(raw[g]) (nearest TF32); -(g+2) lo; g | 2^30 plain; -1 zero."""
    g = tri.gather.astype(np.int64)
    raw = np.asarray(raw, np.float32)
    out = np.zeros(len(g), np.float32)
    plain = (g >= 0) & ((g & (1 << 30)) != 0)
    hi = (g >= 0) & ~plain
    lo = g <= -2
    out[plain] = raw[g[plain] & ~(1 << 30)]
    rnd = lambda v: ((np.ascontiguousarray(v, np.float32).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)
    out[hi] = rnd(raw[g[hi]])
    v = raw[-g[lo] - 2].copy()
    out[lo] = rnd(v - rnd(v))
    return out


def sweep_tri(tri, packed, x, inverse, passes=3):
    """numpy walk of the windowed block-triangular schedule over the packed image of tri_layout.build_tri: tensor memory as
    a NaN-initialised [n, 512] array (unwritten columns must never be read), the scratch area as per-layer [n, K] arrays,
    the chunk stream consumed with a running pointer exactly like the kernel's producer / issuer."""
    from pocomc_b200 import tri_layout as TL
    m = tri.meta
    D, L, T, NB, NW, G, KC = (int(m[i]) for i in (TL.TRI_D, TL.TRI_L, TL.TRI_T, TL.TRI_NB, TL.TRI_NW, TL.TRI_GSIZE, TL.TRI_KCHUNK))
    blocks = m[m[TL.TRI_OFF_BLOCKS]:m[TL.TRI_OFF_BLOCKS] + NB * TL.TB_FIELDS].reshape(NB, -1).astype(np.int64)
    wins = m[m[TL.TRI_OFF_WINDOWS]:m[TL.TRI_OFF_WINDOWS] + NW * TL.TW_FIELDS].reshape(NW, -1).astype(np.int64)
    f32 = np.float32
    trunc = lambda v: (np.ascontiguousarray(v, f32).view(np.uint32) & np.uint32(0xffffe000)).view(f32)     # the tensor core's own cut
    rnd = lambda v: ((np.ascontiguousarray(v, f32).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(f32)
    v = np.array(x, f32, copy=True)
    n = len(v)
    ladj = np.zeros(n, f32)
    log_slope = f32(np.log(1e-3))
    kind = int(m[TL.TRI_KIND])
    PO = TL.tri_out_cols(kind)                    # output columns per order position

    def mma(A, Bh, Bl):
        Ah = rnd(A)
        Al = trunc(A - Ah)
        Dm = Ah.astype(np.float64) @ Bh.T.astype(np.float64)
        if passes > 1:
            Dm += Al.astype(np.float64) @ Bh.T.astype(np.float64) + Ah.astype(np.float64) @ Bl.T.astype(np.float64)
        return Dm.astype(f32)

    for t in (range(T - 1, -1, -1) if inverse else range(T)):
        P = packed[t * tri.tstride:(t + 1) * tri.tstride]
        iperm = np.arange(D) if t % 2 == 0 else D - 1 - np.arange(D)
        acc = np.full((n, 512), np.nan, f32)
        scratch = [np.zeros((n, int(m[TL.TRI_KX_TOTAL])), f32)] + [np.zeros((n, int(m[TL.TRI_KH_TOTAL])), f32) for _ in range(L)]
        cp = int(m[TL.TRI_CHUNK_OFF])                                # running pointer into the chunk stream

        def take_b(N, K):
            nonlocal cp
            hi = P[cp:cp + N * K].reshape(K // 4, N, 4).transpose(1, 0, 2).reshape(N, K)
            lo = P[cp + N * K:cp + 2 * N * K].reshape(K // 4, N, 4).transpose(1, 0, 2).reshape(N, K)
            cp += 2 * N * K
            return hi, lo

        for bi in range(NB):
            k0, nst, NR, W, Kp, win, wc, oc, doff, dn, ks, upd_N, upd_dcol, out_N, out_dcol, flags = (int(q) for q in blocks[bi])
            b0w, nbw, Wp, Op, col_out, Kh, Kx, _ = (int(q) for q in wins[win])
            E = NR - 4
            nrv, q1 = (1 if NR == 4 else 2), (2 * NR + 3) // 4
            first_block = (win == 0 and bi == 0)
            a = [None] + [np.zeros((n, W), f32) for _ in range(L)]
            o = np.zeros((n, PO * G), f32)
            if not first_block:
                for l in range(1, L + 1):
                    a[l] = acc[:, (l - 1) * Wp + wc:(l - 1) * Wp + wc + W].copy()
                o = acc[:, col_out + oc:col_out + oc + PO * G].copy()
                assert np.isfinite(o[:, :PO * nst]).all() and all(np.isfinite(a[l]).all() for l in range(1, L + 1)), (t, bi)
                o = np.nan_to_num(o)
            xb = np.zeros((n, G), f32)
            p = doff

            def take(nf4):
                nonlocal p
                out_ = P[p:p + 4 * nf4]
                p += 4 * nf4
                return out_

            for j in range(G):
                if kind != 0:
                    # spline head: 24 parameter columns; bias 6 f4, then per source group 4 regular units + the extras, 6 f4 each
                    phi = o[:, PO * j:PO * j + PO] + take(6)[None, :]
                    for c in range(j):
                        for src in [4 * c + u_ for u_ in range(4)] + [4 * G + E * c + e_ for e_ in range(E)]:
                            phi = phi + a[L][:, src][:, None] * take(6)[None, :]
                    if j < nst:
                        feat = iperm[k0 + j]
                        res, lj = rqs(phi[:, :23].astype(f32), v[:, feat].copy(), inverse)
                        xk = res if inverse else v[:, feat].copy()
                        ladj = (ladj - lj if inverse else ladj + lj).astype(f32)
                        v[:, feat] = res
                        xb[:, j] = xk
                bo = take(1) if kind == 0 else np.zeros(4, f32)
                shift = o[:, 2 * j] + bo[0]
                sraw = o[:, 2 * j + 1] + bo[1]
                for c in range(j if kind == 0 else 0):
                    for pp in range(2):
                        w4 = take(1)
                        s0, s1 = 4 * c + 2 * pp, 4 * c + 2 * pp + 1
                        shift = shift + w4[0] * a[L][:, s0] + w4[1] * a[L][:, s1]
                        sraw = sraw + w4[2] * a[L][:, s0] + w4[3] * a[L][:, s1]
                    if E == 1:
                        w4 = take(1)
                        shift = shift + w4[0] * a[L][:, 4 * G + c]
                        sraw = sraw + w4[1] * a[L][:, 4 * G + c]
                    elif E == 2:
                        w4 = take(1)
                        e0, e1 = 4 * G + 2 * c, 4 * G + 2 * c + 1
                        shift = shift + w4[0] * a[L][:, e0] + w4[1] * a[L][:, e1]
                        sraw = sraw + w4[2] * a[L][:, e0] + w4[3] * a[L][:, e1]
                if j < nst and kind == 0:
                    feat = iperm[k0 + j]
                    ls = sraw / (f32(1) + np.abs(sraw / log_slope))
                    if inverse:
                        xk = (v[:, feat] - shift) * np.exp(-ls)
                        ladj -= ls
                        v[:, feat] = xk
                    else:
                        xk = v[:, feat].copy()
                        v[:, feat] = xk * np.exp(ls) + shift
                        ladj += ls
                    xb[:, j] = xk
                own = [TL.tri_slot(j, s_, G, E) for s_ in range(NR)]
                pre = a[1][:, own] + take(nrv)[None, :NR]
                for q in range(j // 2 + 1):
                    blk = take(q1)[:2 * NR].reshape(NR, 2)
                    pre = pre + xb[:, 2 * q][:, None] * blk[None, :, 0] + xb[:, 2 * q + 1][:, None] * blk[None, :, 1]
                a[1][:, own] = np.maximum(pre, 0)
                for l in range(2, L + 1):
                    pre = a[l][:, own] + take(nrv)[None, :NR]
                    for c in range(j + 1):
                        blk = take(NR).reshape(2, NR, 2)
                        for pp in range(2):
                            for h in range(2):
                                pre = pre + a[l - 1][:, 4 * c + 2 * pp + h][:, None] * blk[None, pp, :, h]
                        if E == 1:
                            blk = take(2)[:NR]
                            pre = pre + a[l - 1][:, 4 * G + c][:, None] * blk[None, :]
                        elif E == 2:
                            blk = take(3).reshape(NR, 2)
                            pre = pre + a[l - 1][:, 4 * G + 2 * c][:, None] * blk[None, :, 0] + a[l - 1][:, 4 * G + 2 * c + 1][:, None] * blk[None, :, 1]
                    a[l][:, own] = np.maximum(pre + a[l - 1][:, own], 0)
            assert p == doff + dn
            # the block's activations: A tiles (K padded to a multiple of 8) and, when another window follows, the scratch area
            Ax = np.zeros((n, 8), f32); Ax[:, :G] = xb
            At = [Ax]
            for l in range(1, L + 1):
                Ap = np.zeros((n, Kp), f32); Ap[:, :W] = a[l]
                At.append(Ap)
            if NW > 1:
                scratch[0][:, 8 * bi:8 * bi + 8] = Ax
                for l in range(1, L + 1):
                    scratch[l][:, ks:ks + Kp] = At[l]
            if not (flags & 1):
                first = first_block
                for op in (1, 2, 3, 4):
                    N, dcol = (upd_N, (op - 1) * Wp + upd_dcol) if op < 4 else (out_N, col_out + out_dcol)
                    A = At[op - 1]
                    Bh, Bl = take_b(N, A.shape[1])
                    Dm = mma(A, Bh, Bl)
                    if first:
                        acc[:, dcol:dcol + N] = Dm
                    else:
                        acc[:, dcol:dcol + N] += Dm
            elif not (flags & 2):
                nb0, nnb, nWp, nOp, ncol_out, nKh, nKx, _ = (int(q) for q in wins[win + 1])
                for op in (1, 2, 3, 4):
                    ktot = nKx if op == 1 else nKh
                    N, dcol = (nWp, (op - 1) * nWp) if op < 4 else (nOp, ncol_out)
                    kc_ = KC if op < 4 else int(m[TL.TRI_KCHUNK_OUT])
                    for k_ in range(0, ktot, kc_):
                        ke = min(kc_, ktot - k_)
                        A = scratch[op - 1][:, k_:k_ + ke]
                        Bh, Bl = take_b(N, ke)
                        Dm = mma(A, Bh, Bl)
                        if k_ == 0:
                            acc[:, dcol:dcol + N] = Dm
                        else:
                            acc[:, dcol:dcol + N] += Dm
        assert cp == tri.tstride, (cp, tri.tstride)
    return v, ladj
