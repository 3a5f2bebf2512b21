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


def _mma_dense(w, pos, K, n_cols):
    """Inverse of made_layout._mma_slab: (dense [K8, NT*8] matrix, floats consumed)."""
    K8, NT = (K + 7) // 8 * 8, (n_cols + 7) // 8
    cnt = NT * (K8 // 8) * 64
    blk = w[pos:pos + cnt].reshape(NT, K8 // 8, 32, 2)
    dense = np.zeros((K8, NT * 8), np.float32)
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    for nt in range(NT):
        for ks in range(K8 // 8):
            for j in range(2):
                dense[8 * ks + t + 4 * j, 8 * nt + g] = blk[nt, ks, :, j]
    return dense, cnt


def sweep_stream_mma(layout, stream, packed, x, inverse):
    """Emulates the warp-MMA stream kernel's walk over the "mma" stream variant (fp32 arithmetic)."""
    m = stream.meta
    D, H, L, T, ng, total = (int(m[i]) for i in (ML.M_D, ML.M_H, ML.M_L, ML.M_T, ML.M_NG, ML.M_TOTAL))
    gstart = m[m[ML.M_OFF_GSTART]:m[ML.M_OFF_GSTART] + ng + 1]
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
                dense, c = _mma_dense(w, pos, ek, total); pos += c
                assert not dense[ek:].any() and not dense[:, total:].any()
                ntp = (total + 7) // 8 * 8
                phi = act[L - 1][:, :ek] @ dense[:ek] + w[pos:pos + ntp]; pos += ntp
                v = cur[:, feat].copy()
                res, l = uni(phi[:, :layout.tp], v, inverse)
                ladj = ladj - l if inverse else ladj + l
                xs[:, k] = res if inverse else v
                cur[:, feat] = res
                g = k + 1
                if g > ng or gstart[g] == gstart[g - 1]:
                    continue
                gs, ge = int(gstart[g - 1]), int(gstart[g])
                for l_ in range(L):
                    nrows = g if l_ == 0 else ge
                    src = xs[:, :g] if l_ == 0 else act[l_ - 1][:, :ge]
                    dense, c = _mma_dense(w, pos, nrows, ge - gs); pos += c
                    n8 = (ge - gs + 7) // 8 * 8
                    pre = src @ dense[:nrows] + w[pos:pos + n8]; pos += n8
                    pre = pre[:, :ge - gs]
                    if l_ > 0:
                        pre = pre + act[l_ - 1][:, gs:ge]
                    act[l_][:, gs:ge] = np.maximum(pre, 0)
            assert pos == cnt
    return cur, ladj


# ---------------------------------------------------------------------------------------------
# blocked sweep (csrc/flow_block.cu): interpreter over made_layout.build_block's program
# ---------------------------------------------------------------------------------------------
def pack_block(block, raw):
    """pmc_flow_tc_pack's codes: >= 0 hi(raw[g]) (13 low mantissa bits cleared), -(g+2) lo = raw - hi,
    g | 2^30 plain copy, -1 zero."""
    raw = np.asarray(raw, np.float32)
    g = block.gather.astype(np.int64)
    out = np.zeros(g.size, np.float32)
    plain = g >= ML.BLOCK_PLAIN
    out[plain] = raw[g[plain] - ML.BLOCK_PLAIN]
    hi_all = (raw.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    hi = (g >= 0) & ~plain
    out[hi] = hi_all[g[hi]]
    lo = g <= -2
    idx = -g[lo] - 2
    out[lo] = raw[idx] - hi_all[idx]
    return out


def sweep_block(block, packed, x, inverse):
    """Executes the op program exactly like the kernel's consumer warp (8 particles at a time), with
    the kernel's array strides and fragment orders; arithmetic in fp32/fp64 numpy."""
    m = block.meta
    D, L, T = (int(m[i]) for i in (ML.M_D, ML.M_L, ML.M_T))
    nops, sx, so = int(m[ML.M_NOPS]), int(m[ML.M_SX]), int(m[ML.M_SO])
    sh = int(m[ML.M_HPB]) + 4
    prog = m[m[ML.M_OFF_PROG]:m[ML.M_OFF_PROG] + 8 * nops].reshape(nops, 8)
    chunks = m[m[ML.M_OFF_CHUNKS]:m[ML.M_OFF_CHUNKS] + 4 * m[ML.M_NCHUNKS]].reshape(-1, 4)
    lane = np.arange(32)
    fr, fc = lane >> 2, lane & 3
    x = np.asarray(x, np.float32)
    n_all = len(x)
    out_all = np.zeros_like(x)
    ladj_all = np.zeros(n_all, np.float32)
    for row0 in range(0, n_all, 8):
        rows = min(8, n_all - row0)
        cur = np.zeros((D, 8), np.float32)
        cur[:, :rows] = x[row0:row0 + rows].T
        xs = np.zeros((8, sx), np.float32)
        act = np.zeros((L, 8, sh), np.float32)
        ph = np.zeros((8, so), np.float32)
        arrays = [xs] + [act[l] for l in range(L)] + [ph]
        ladj = np.zeros(8, np.float32)
        for tt in range(T):
            t = T - 1 - tt if inverse else tt
            ci = -1
            w, pos = None, 0
            acc = None
            onew = np.zeros((8, 2))
            for op in prog:
                typ, a, b, c, d, e, _, flags = (int(v) for v in op)
                if flags & ML.BF_NEWCHUNK:
                    assert w is None or pos == len(w), (pos, len(w))
                    ci += 1
                    off, cnt = int(chunks[ci, 2]), int(chunks[ci, 3])
                    w = packed[t * block.tstride + off: t * block.tstride + off + cnt]
                    pos = 0
                if typ == ML.OP_MMA:
                    nt = int(op[6])
                    if flags & ML.BF_FIRST:
                        acc = w[pos:pos + 128 * nt].reshape(nt, 32, 4).astype(np.float64); pos += 128 * nt
                    src = arrays[a]
                    for ks in range(c):
                        B = src[:, b + 8 * ks: b + 8 * ks + 8].T.astype(np.float64)        # [k][particle]
                        for tl in range(nt):
                            ah = w[pos:pos + 128].reshape(32, 4); al = w[pos + 128:pos + 256].reshape(32, 4); pos += 256
                            A = np.zeros((16, 8), np.float64)
                            A[fr, fc] = ah[:, 0].astype(np.float64) + al[:, 0]; A[fr + 8, fc] = ah[:, 1].astype(np.float64) + al[:, 1]
                            A[fr, fc + 4] = ah[:, 2].astype(np.float64) + al[:, 2]; A[fr + 8, fc + 4] = ah[:, 3].astype(np.float64) + al[:, 3]
                            C = A @ B                                                       # [16 units][8 particles]
                            acc[tl, :, 0] += C[fr, 2 * fc]; acc[tl, :, 1] += C[fr, 2 * fc + 1]
                            acc[tl, :, 2] += C[fr + 8, 2 * fc]; acc[tl, :, 3] += C[fr + 8, 2 * fc + 1]
                    if flags & ML.BF_LAST:
                        dst = arrays[d]
                        for tl in range(nt):
                            r0 = e + 16 * tl
                            dst[2 * fc, r0 + fr] = acc[tl, :, 0]; dst[2 * fc + 1, r0 + fr] = acc[tl, :, 1]
                            dst[2 * fc, r0 + fr + 8] = acc[tl, :, 2]; dst[2 * fc + 1, r0 + fr + 8] = acc[tl, :, 3]
                elif typ == ML.OP_STEP:
                    k, c_ = a & 0xFFFF, a >> 16
                    out_old, pbj = b & 0xFFFF, b >> 16
                    u0, cnt_u = c & 0xFFFF, c >> 16
                    l0_d0, l0_r = d & 0xFFFF, d >> 16
                    lh = e
                    P4 = (cnt_u + 3) // 4 * 4

                    def take(nfl):
                        nonlocal pos
                        v = w[pos:pos + nfl].astype(np.float64); pos += nfl
                        return v

                    def old_dot(src, r0, nrows, slots):
                        wv = take(nrows * slots).reshape(nrows // 4, slots, 4)
                        xv = src[:, r0:r0 + nrows].reshape(8, nrows // 4, 4).astype(np.float64)
                        return np.einsum("pgr,gsr->ps", xv, wv)

                    if flags & ML.BF_BLOCKFIRST:
                        onew = np.zeros((8, 2))
                    phi = ph[:, c_:c_ + 2].astype(np.float64) + onew
                    if out_old:
                        phi = phi + old_dot(act[L - 1], pbj, out_old, 2)
                    feat = D - 1 - k if t % 2 else k
                    v = cur[feat].copy()
                    res, l = affine(phi.astype(np.float32), v, inverse)
                    ladj = (ladj - l if inverse else ladj + l).astype(np.float32)
                    xk = res if inverse else v
                    xs[:, k] = xk
                    cur[feat] = res
                    onew = np.zeros((8, 2))
                    if cnt_u:
                        pre = act[0][:, u0:u0 + P4].astype(np.float64)
                        if l0_r:
                            pre = pre + old_dot(xs, l0_d0, l0_r, P4)
                        pre = pre + xk[:, None].astype(np.float64) * take(P4)[None, :]
                        h = np.maximum(pre, 0).astype(np.float32)
                        act[0][:, u0:u0 + P4] = h
                        for l_ in range(1, L):
                            pre = act[l_][:, u0:u0 + P4].astype(np.float64) + h
                            if lh:
                                pre = pre + old_dot(act[l_ - 1], pbj, lh, P4)
                            pre = pre + h.astype(np.float64) @ take(P4 * P4).reshape(P4, P4).T
                            h = np.maximum(pre, 0).astype(np.float32)
                            act[l_][:, u0:u0 + P4] = h
                        if flags & ML.BF_NEXTOUT:
                            onew = h.astype(np.float64) @ take(2 * P4).reshape(2, P4).T
                else:
                    raise AssertionError(typ)
            assert pos == len(w) and ci == len(chunks) - 1
        out_all[row0:row0 + rows] = cur[:, :rows].T
        ladj_all[row0:row0 + rows] = ladj[:rows]
    return out_all, ladj_all


def sweep_stream_tip(layout, stream, packed, x, inverse):
    """Emulates the bulk/tip schedule of csrc/flow_tip.cu over made_layout.build_stream_tip's stream: every dot
    product = bulk (inputs finished one order position earlier, computed ahead) + tip (the degree group born in
    this position).  Walks the stream with a running float4 offset exactly like the kernel."""
    m = stream.meta
    assert int(m[ML.M_VERSION]) == 5
    D, H, L, T, ng = (int(m[i]) for i in (ML.M_D, ML.M_H, ML.M_L, ML.M_T, ML.M_NG))
    gstart = m[m[ML.M_OFF_GSTART]:m[ML.M_OFF_GSTART] + ng + 1].astype(np.int64)
    nchunk = m[m[ML.M_OFF_NCHUNK]:m[ML.M_OFF_NCHUNK] + ng].astype(np.int64)
    chunks = m[m[ML.M_OFF_CHUNKS]:m[ML.M_OFF_CHUNKS] + 4 * m[ML.M_NCHUNKS]].reshape(-1, 4)
    p16 = lambda v: (int(v) + 15) // 16 * 16
    cur = np.array(x, np.float32, copy=True)
    n = len(cur)
    ladj = np.zeros(n, np.float32)
    for tt in range(T):
        t = T - 1 - tt if inverse else tt
        xs = np.zeros((n, D), np.float32)
        act = np.zeros((L, n, H), np.float32)
        bout = np.zeros((n, 4), np.float32)                     # bulk part of phi for the CURRENT position
        fresh = np.zeros((n, 0), np.float32)                    # last-layer activations of the group born one position ago
        for k0, k1, off, cnt in chunks:
            w = packed[t * stream.tstride + off: t * stream.tstride + off + cnt].reshape(-1, 4)
            pos = 0
            for k in range(k0, k1):
                feat = D - 1 - k if t % 2 else k
                # ---- tip head: phi = bulk + out tip + bias -> univariate map
                nchp = int(nchunk[k - 1]) if k >= 1 else 0
                assert fresh.shape[1] == 4 * nchp
                phi = bout + fresh @ w[pos:pos + 4 * nchp] + w[pos + 4 * nchp]
                pos += 4 * nchp + 1
                v = cur[:, feat].copy()
                res, l = affine(phi, v, inverse)
                ladj = ladj - l if inverse else ladj + l
                xk = res if inverse else v
                xs[:, k] = xk
                cur[:, feat] = res
                g = k + 1
                ek = int(gstart[k])
                bulk = None
                if g <= ng:
                    nch = int(nchunk[g - 1])
                    gs, ge = int(gstart[g - 1]), int(gstart[g])
                    bulk = np.zeros((L, n, 4 * nch), np.float32)
                    for l_ in range(L):
                        nrows = k if l_ == 0 else ek
                        src = xs[:, :k] if l_ == 0 else act[l_ - 1][:, :ek]
                        for c in range(nch):
                            slab = w[pos:pos + p16(nrows)]; pos += p16(nrows)
                            assert not slab[nrows:].any()
                            bulk[l_][:, 4 * c:4 * c + 4] = src @ slab[:nrows]
                if k + 1 < D:
                    slab = w[pos:pos + p16(ek)]; pos += p16(ek)
                    assert not slab[ek:].any()
                    bout = act[L - 1][:, :ek] @ slab[:ek]
                if g <= ng:
                    prev = None                                     # fresh activations of the previous layer [n, 4 nch]
                    for l_ in range(L):
                        new = np.zeros((n, 4 * nch), np.float32)
                        for c in range(nch):
                            for q in range(4):
                                j = 4 * c + q
                                head = w[pos]; pos += 1
                                if l_ == 0:
                                    pre = bulk[0][:, j] + head[0] + head[1] * xk
                                else:
                                    tipw = w[pos:pos + nch].reshape(-1); pos += nch
                                    pre = bulk[l_][:, j] + head[0] + prev @ tipw + prev[:, j]      # + residual
                                new[:, j] = np.maximum(pre, 0)
                        real = ge - gs
                        assert not new[:, real:].any()              # padding units stay exactly zero
                        act[l_][:, gs:ge] = new[:, :real]
                        prev = new
                    fresh = prev
            assert pos == cnt // 4, (pos, cnt)
    return cur, ladj


def sweep_tip_lanes(stream, packed, x, inverse, ppl=1):
    """Lane-level transliteration of csrc/flow_tip.cu (made_sweep_tip_kernel for ppl = 1, made_sweep_tip_ppl_kernel
    for ppl = 2 | 4): one warp = 32 lanes, lane = q * 8 + p carrying particles p * ppl + e; shared memory as flat
    float arrays with the kernel's index expressions, shuffles as lane permutations.  Pins the lane mapping
    (activation addressing of the blocked dot products, reduce-scatter, shuffle exchange, guarded writes, stream
    pointer arithmetic) on the CPU; x may hold any number of rows (processed warp tile by warp tile)."""
    m = stream.meta
    assert int(m[ML.M_VERSION]) == 5
    D, H, L, T, ng = (int(m[i]) for i in (ML.M_D, ML.M_H, ML.M_L, ML.M_T, ML.M_NG))
    Dp, Hp = (D + 15) // 16 * 16, (H + 15) // 16 * 16
    gstart = m[m[ML.M_OFF_GSTART]:m[ML.M_OFF_GSTART] + ng + 1].astype(np.int64)
    nchunk = m[m[ML.M_OFF_NCHUNK]:m[ML.M_OFF_NCHUNK] + ng].astype(np.int64)
    chunks = m[m[ML.M_OFF_CHUNKS]:m[ML.M_OFF_CHUNKS] + 4 * m[ML.M_NCHUNKS]].reshape(-1, 4)
    maxch = int(m[ML.M_MAXCH])
    PWV = 8 * ppl
    lane = np.arange(32)
    p_, q_ = lane & 7, lane >> 3
    pe = p_ * ppl
    lane_off = ppl * lane
    f32 = np.float32
    LS = f32(math.log(1e-3))
    x = np.asarray(x, f32)
    n = len(x)
    out = np.zeros_like(x)
    ladj_out = np.zeros(n, f32)

    def shfl(v, src):                      # __shfl_sync(FULL, v, src): v [32] per-lane values
        return v[src]

    def dot4(w4, wpos, src, base, rows):
        """dot4_partial(_ppl): acc[lane, e, 4]"""
        acc = np.zeros((32, ppl, 4), f32)
        wp, ap = wpos + q_, base + lane_off
        for _ in range(0, rows, 16):
            for j in range(4):
                w = w4[wp + 4 * j]                                        # [32, 4]
                for e in range(ppl):
                    xv = src[ap + 32 * ppl * j + e]
                    acc[:, e, :] += w * xv[:, None]
            wp = wp + 16
            ap = ap + 128 * ppl
        return acc

    def reduce_scatter4(a):                # a [32, 4] -> [32]: lane q ends with unit q's total
        hi = (lane & 16) != 0
        k0 = np.where(hi, a[:, 2], a[:, 0]) + shfl(np.where(hi, a[:, 0], a[:, 2]), lane ^ 16)
        k1 = np.where(hi, a[:, 3], a[:, 1]) + shfl(np.where(hi, a[:, 1], a[:, 3]), lane ^ 16)
        mid = (lane & 8) != 0
        return np.where(mid, k1, k0) + shfl(np.where(mid, k0, k1), lane ^ 8)

    for row0 in range(0, n, PWV):
        rows = min(PWV, n - row0)
        cur = np.zeros(D * PWV, f32)
        xs = np.zeros(Dp * PWV, f32)
        act = np.zeros(L * Hp * PWV, f32)
        for r in range(rows):
            for c in range(D):
                cur[c * PWV + r] = x[row0 + r, c]
        ladj = np.zeros((32, ppl), f32)
        for tt in range(T):
            t = T - 1 - tt if inverse else tt
            rev = bool(t & 1)
            bout = np.zeros((32, ppl, 2), f32)
            fresh = np.zeros((32, 4 * maxch, ppl), f32)
            for k0, k1, off, cnt in chunks:
                w4 = packed[t * stream.tstride + off: t * stream.tstride + off + cnt].reshape(-1, 4)
                w = 0
                for k in range(k0, k1):
                    feat = D - 1 - k if rev else k
                    nchp = int(nchunk[k - 1]) if k >= 1 else 0
                    phi = bout.copy()
                    for cc in range(nchp):
                        for j in range(4):
                            t4 = w4[w + 4 * cc + j]
                            for e in range(ppl):
                                phi[:, e, 0] += fresh[:, 4 * cc + j, e] * t4[0]
                                phi[:, e, 1] += fresh[:, 4 * cc + j, e] * t4[1]
                    b4 = w4[w + 4 * nchp]
                    w += 4 * nchp + 1
                    xk = np.zeros((32, ppl), f32)
                    res = np.zeros((32, ppl), f32)
                    for e in range(ppl):
                        s0, s1 = phi[:, e, 0] + b4[0], phi[:, e, 1] + b4[1]
                        v = cur[feat * PWV + pe + e]
                        ls = (s1 / (f32(1) + np.abs(s1 / LS))).astype(f32)
                        sc = np.exp(ls).astype(f32)
                        res[:, e] = (v - s0) / sc if inverse else v * sc + s0
                        ladj[:, e] = ladj[:, e] - ls if inverse else ladj[:, e] + ls
                        xk[:, e] = res[:, e] if inverse else v
                    g = k + 1
                    has_group = g <= ng
                    nch = int(nchunk[k]) if has_group else 0
                    ek16 = (int(gstart[k]) + 15) // 16 * 16
                    k16 = (k + 15) // 16 * 16
                    bulk = np.zeros((L, maxch, 32, ppl), f32)
                    if has_group:
                        for l_ in range(L):
                            nrows = k16 if l_ == 0 else ek16
                            src, base = (xs, 0) if l_ == 0 else (act, (l_ - 1) * Hp * PWV)
                            for cc in range(nch):
                                acc = dot4(w4, w, src, base, nrows)
                                w += nrows
                                for e in range(ppl):
                                    bulk[l_, cc, :, e] = reduce_scatter4(acc[:, e, :])
                    nb = np.zeros((32, ppl, 2), f32)
                    if k + 1 < D:
                        acc = dot4(w4, w, act, (L - 1) * Hp * PWV, ek16)
                        w += ek16
                        for e in range(ppl):
                            for o in range(2):
                                a = acc[:, e, o] + shfl(acc[:, e, o], lane ^ 8)
                                nb[:, e, o] = a + shfl(a, lane ^ 16)
                    for ln in np.nonzero(q_ == 0)[0]:
                        for e in range(ppl):
                            xs[k * PWV + pe[ln] + e] = xk[ln, e]
                            cur[feat * PWV + pe[ln] + e] = res[ln, e]
                    if has_group:
                        gs, gsz = int(gstart[k]), int(gstart[k + 1] - gstart[k])
                        mine = np.zeros((maxch, 32, ppl), f32)
                        prev = np.zeros((4 * maxch, 32, ppl), f32)
                        for l_ in range(L):
                            stride = 1 if l_ == 0 else 1 + nch
                            nw = np.zeros((maxch, 32, ppl), f32)
                            for cc in range(nch):
                                base = w + (4 * cc + q_) * stride                   # per lane
                                head = w4[base]
                                for e in range(ppl):
                                    pre = bulk[l_, cc, :, e] + head[:, 0]
                                    if l_ == 0:
                                        pre = pre + head[:, 1] * xk[:, e]
                                    else:
                                        for c2 in range(nch):
                                            t4 = w4[base + 1 + c2]
                                            for jj in range(4):
                                                pre = pre + t4[:, jj] * prev[4 * c2 + jj, :, e]
                                        pre = pre + mine[cc, :, e]
                                    nw[cc, :, e] = np.maximum(pre, 0)
                                    for ln in range(32):
                                        if 4 * cc + q_[ln] < gsz:
                                            act[l_ * Hp * PWV + (gs + 4 * cc + q_[ln]) * PWV + pe[ln] + e] = nw[cc, ln, e]
                            w += 4 * nch * stride
                            for cc in range(maxch):
                                for e in range(ppl):
                                    mine[cc, :, e] = nw[cc, :, e]
                                    for j in range(4):
                                        prev[4 * cc + j, :, e] = shfl(nw[cc, :, e], 8 * j + p_)
                        for j in range(4 * maxch):
                            fresh[:, j, :] = prev[j]
                    bout = nb
                assert w == cnt // 4
        for r in range(rows):
            for c in range(D):
                out[row0 + r, c] = cur[c * PWV + r]
        for ln in np.nonzero(q_ == 0)[0]:
            for e in range(ppl):
                if pe[ln] + e < rows:
                    ladj_out[row0 + pe[ln] + e] = ladj[ln, e]
    return out, ladj_out
