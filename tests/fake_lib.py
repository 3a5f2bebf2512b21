"""numpy stand-ins for the persistent-sampling entry points of libpmc_b200 (test infrastructure only).

Lets the CPU suite exercise the HOST logic that sits on top of those kernels -- in particular the multi-rank
exchange of pocomc_b200.sharded over gloo -- where no GPU exists.  ``install(monkeypatch)`` redirects
``pocomc_b200._lib.call`` for the names below, makes ``_lib.device()`` the CPU and leaves everything else alone;
the arithmetic follows csrc/smc_ops.cu (which the GPU tests pin against the reference's goldens)."""
import ctypes as C
import math

import numpy as np
import torch


def _arr(p, n, dtype=np.float64):
    addr = p.value if isinstance(p, C.c_void_p) else int(p)
    ctype = {np.float64: C.c_double, np.int64: C.c_int64}[dtype]
    return np.ctypeslib.as_array((ctype * int(n)).from_address(addr))


def ps_append(logl, den, beta, logz, t_new, t_total, n):
    """csrc/smc_ops.cu ps_append_kernel: den[t, j] = logaddexp_k(beta_k logl[t, j] - logz_k), folded incrementally"""
    L = _arr(logl, t_total * n).reshape(t_total, n)
    Dn = _arr(den, t_total * n).reshape(t_total, n)
    b, z = _arr(beta, t_total), _arr(logz, t_total)
    for t in range(t_total):
        if t < t_new:
            acc, start = Dn[t].copy(), t_new
        else:
            acc, start = L[t] * b[0] - z[0], 1
        for k in range(start, t_total):
            acc = np.logaddexp(acc, L[t] * b[k] - z[k])
        Dn[t] = acc


def _logw(logl, den, beta_f, t_total, n):
    m = t_total * n
    return _arr(logl, m) * beta_f - (_arr(den, m) - math.log(t_total))


def ps_reduce(logl, den, beta_f, t_total, n, uss_k, scratch, out4):
    lw = _logw(logl, den, beta_f, t_total, n)
    mx = lw.max()
    e = np.exp(lw - mx)
    o = _arr(out4, 4)
    o[0], o[1], o[2] = mx, e.sum(), (e * e).sum()
    o[3] = np.sum(1.0 - (1.0 - e / e.sum()) ** uss_k) if uss_k > 0 else 0.0


def ps_weights(logl, den, beta_f, t_total, n, stats, w, logw):
    lw = _logw(logl, den, beta_f, t_total, n)
    st = _arr(stats, 4)
    if w is not None and getattr(w, "value", w):
        _arr(w, t_total * n)[:] = np.exp(lw - st[0]) / st[1]
    if logw is not None and getattr(logw, "value", logw):
        _arr(logw, t_total * n)[:] = lw - (st[0] + math.log(st[1]))


def weight_stats(w, m, uss_k, scratch, out3):
    """tools.py:56-93 on an (unnormalised) weight vector: [sum w, sum w^2, sum 1 - (1 - w / sum)^k]"""
    x = _arr(w, m)
    o = _arr(out3, 3)
    o[0], o[1] = x.sum(), (x * x).sum()
    o[2] = np.sum(1.0 - (1.0 - x / x.sum()) ** uss_k) if uss_k > 0 else 0.0


def trim_threshold(ws, m, ess_frac, bins, scratch, out3):
    """tools.py:10-53 on the ascending-sorted weights: walk the percentile grid linspace(0, 99, bins) from the top
    until ESS(w[w >= thr]) / ESS(w) >= ess_frac; out3 = [threshold, kept sum, grid index]"""
    x = _arr(ws, m)
    total = x.sum()
    ess_total = total * total / (x * x).sum()
    grid = np.linspace(0.0, 99.0, int(bins))
    i = int(bins) - 1
    while True:
        thr = np.percentile(x, grid[i])
        kept = x[x >= thr]
        if kept.sum() ** 2 / (kept * kept).sum() / ess_total >= ess_frac or i == 0:
            break
        i -= 1
    o = _arr(out3, 3)
    o[0], o[1], o[2] = thr, kept.sum(), float(i)


TABLE = {"pmc_ps_append": ps_append, "pmc_ps_reduce": ps_reduce, "pmc_ps_weights": ps_weights,
         "pmc_weight_stats": weight_stats, "pmc_trim_threshold": trim_threshold}


def install(monkeypatch):
    from pocomc_b200 import _lib

    def call(name, *args):
        if name not in TABLE:
            raise RuntimeError(f"fake_lib: no numpy stand-in for {name}")
        TABLE[name](*args)

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(_lib, "require_cuda", lambda: None)
    monkeypatch.setattr(_lib, "device", lambda: torch.device("cpu"))
