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
