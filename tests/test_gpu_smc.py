"""GPU parity: persistent-sampling weights, ESS/USS, trimming, resampling and evidence reductions
against vectors recorded from the reference (particles.py / tools.py / sampler.py).
log-weights / logZ: 1e-12 relative (north_star asks 1e-5); resampling + trimming indices bit-exact."""
import numpy as np
import pytest
import torch

import smc_ref as O

pytestmark = pytest.mark.gpu


def _particles(g, upto=None):
    from pocomc_b200.particles import Particles
    p = Particles(g["logl"].shape[1], 3)
    for t in range(len(g["beta"]) if upto is None else upto):
        p.update(dict(logl=g["logl"][t], beta=g["beta"][t], logz=g["logz"][t]))
    return p


def test_compute_logw_and_logz(golden):
    g = golden("smc")
    p = _particles(g)
    for tag, b, nz in (("b1", 1.0, True), ("b05", 0.5, True), ("b03raw", 0.3, False), ("b0", 0.0, True)):
        lw, lz = p.compute_logw_and_logz(b, normalize=nz)
        np.testing.assert_allclose(lw, g[f"logw_{tag}"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(lz, g[f"logz_{tag}"], rtol=1e-12)


def test_incremental_append_equals_rebuild(golden):
    """folding iterations in one at a time gives the same denominator as one big append"""
    g = golden("smc")
    p = _particles(g, upto=3)
    p.compute_logw_and_logz(0.5)
    for t in range(3, len(g["beta"])):
        p.update(dict(logl=g["logl"][t], beta=g["beta"][t], logz=g["logz"][t]))
        if t % 2:
            p.compute_logw_and_logz(0.5)
    lw, lz = p.compute_logw_and_logz(0.5)
    np.testing.assert_allclose(lw, g["logw_b05"], rtol=1e-12, atol=1e-12)
    pr = p.probe(0.5, uss_k=100)
    np.testing.assert_allclose(pr["ess"], g["ess_b05"], rtol=1e-11)
    np.testing.assert_allclose(pr["uss"], g["uss_b05_k100"], rtol=1e-11)
    # KAT
    from pocomc_b200.particles import Particles
    q = Particles(2, 1)
    q.update(dict(logl=np.array([-1., -2]), beta=0.0, logz=0.0))
    q.update(dict(logl=np.array([-0.5, -3]), beta=0.5, logz=-0.7))
    lw, lz = q.compute_logw_and_logz(1.0)
    np.testing.assert_allclose(lw, [-1.132334848400722, -1.8885512234876576, -0.7774449250165852, -2.7052966449669085], rtol=1e-13)
    np.testing.assert_allclose(lz, -1.358951201540815, rtol=1e-13)
    lw, lz = q.compute_logw_and_logz(0.5, normalize=False)
    np.testing.assert_allclose(lw, [-0.6049916888216466, -0.8612080639085818, -0.5001017654375096, -1.1779534853878324], rtol=1e-13)
    np.testing.assert_allclose(lz, -0.753371118241601, rtol=1e-13)


def test_tools_match_reference(golden):
    from pocomc_b200 import tools
    g = golden("smc")
    lw = g["logw_b05"]
    w = np.exp(lw - lw.max())
    np.testing.assert_allclose(tools.effective_sample_size(w.copy()), g["ess_b05"], rtol=1e-12)
    np.testing.assert_allclose(tools.unique_sample_size(w.copy()), g["uss_b05"], rtol=1e-12)
    np.testing.assert_allclose(tools.unique_sample_size(w.copy(), k=100), g["uss_b05_k100"], rtol=1e-12)
    np.testing.assert_allclose(tools.compute_ess(lw), g["compute_ess_b05"], rtol=1e-12)
    np.testing.assert_allclose(tools.increment_logz(lw), g["increment_logz_b05"], rtol=1e-12, atol=1e-13)
    assert tools.compute_ess(np.array([0.3])) == 1.0                      # reference tests/test_tools.py
    np.testing.assert_allclose(tools.effective_sample_size(np.array([1., 2, 3, 4])), 3.333333333333333, rtol=1e-14)
    np.testing.assert_allclose(tools.unique_sample_size(np.ones(512), k=256), 201.60809550983944, rtol=1e-13)
    wn = w / w.sum()
    for tag, e, b in (("a", 0.99, 1000), ("b", 0.9, 50), ("c", 0.999, 200)):
        idx, wt = tools.trim_weights(np.arange(len(wn)), wn.copy(), ess=e, bins=b)
        np.testing.assert_array_equal(idx, g[f"trim_{tag}_idx"])
        np.testing.assert_allclose(wt, g[f"trim_{tag}_w"], rtol=1e-13)
    idx, wt = tools.trim_weights(np.arange(10), np.arange(1., 11.), ess=0.9, bins=10)
    np.testing.assert_array_equal(idx, np.arange(2, 10))
    # resampling: bit-exact indices
    np.random.seed(123)
    np.testing.assert_array_equal(tools.systematic_resample(500, wn.copy()), g["syst_idx"])
    np.random.seed(0)
    np.testing.assert_array_equal(tools.systematic_resample(4, np.array([0.6, 0.2, 0.15, 0.05])), [0, 0, 1, 2])
    dev = torch.device("cuda")
    idx = tools.multinomial_resample_device(torch.from_numpy(wn).to(dev), torch.from_numpy(g["mult_r"]).to(dev))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["mult_idx"])
    cdf = tools.cumsum_device(torch.from_numpy(wn).to(dev)).cpu().numpy()
    np.testing.assert_array_equal(cdf, np.cumsum(wn))                    # same sequential f64 order
    rows = tools.gather_rows_device(torch.from_numpy(g["logl"].T.copy()).to(dev), idx[:50] % g["logl"].shape[1])
    np.testing.assert_array_equal(rows.cpu().numpy(), g["logl"].T[(g["mult_idx"][:50] % g["logl"].shape[1])])


def test_large_trim_and_resample_properties():
    """full-size (1e6) checks through size-independent properties + the O(n log n) oracle"""
    from pocomc_b200 import tools
    rng = np.random.default_rng(0)
    w = rng.pareto(1.5, size=1_000_000)
    w /= w.sum()
    dev = torch.device("cuda")
    keep, wt = tools.trim_weights_device(torch.from_numpy(w).to(dev), 0.99, 1000)
    keep_ref, wt_ref, _ = O.trim_weights_sorted(w.copy(), 0.99, 1000)
    np.testing.assert_array_equal(keep.cpu().numpy(), keep_ref)
    np.testing.assert_allclose(wt.cpu().numpy(), wt_ref, rtol=1e-12)
    r = rng.random(100_000)
    idx = tools.multinomial_resample_device(torch.from_numpy(w).to(dev), torch.from_numpy(r).to(dev)).cpu().numpy()
    np.testing.assert_array_equal(idx, O.multinomial_resample(w, r))
    idx = tools.systematic_resample_device(50_000, torch.from_numpy(w).to(dev), 0.37).cpu().numpy()
    assert np.all(np.diff(idx) >= 0)
    np.testing.assert_array_equal(idx, np.searchsorted(np.cumsum(w), (0.37 + np.arange(50_000)) / 50_000, side="left"))


def test_evidence_reductions():
    from pocomc_b200 import _lib
    rng = np.random.default_rng(1)
    n, nb = 4096, 300
    logw = rng.normal(size=n) * 3 - 40
    logw[5] = -np.inf
    boot = rng.integers(0, n, size=(nb, n))
    lz_ref, err_ref = O.flow_is_evidence(logw, 0, 0, 0, boot)
    dev = torch.device("cuda")
    lw = torch.from_numpy(logw).to(dev)
    scratch = torch.empty(int(_lib.load().pmc_ps_scratch_size(n)), dtype=torch.float64, device=dev)
    out2 = torch.empty(2, dtype=torch.float64, device=dev)
    _lib.call("pmc_lse", _lib.ptr(lw), n, _lib.ptr(scratch), _lib.ptr(out2))
    np.testing.assert_allclose(out2[0].item(), lz_ref, rtol=1e-13)
    out = torch.empty(nb, dtype=torch.float64, device=dev)
    _lib.call("pmc_lse_bootstrap", _lib.ptr(lw), _lib.ptr(torch.from_numpy(boot).to(dev)), n, nb, _lib.ptr(out))
    np.testing.assert_allclose(np.std(out.cpu().numpy()), err_ref, rtol=1e-10)


def test_bootstrap_on_device_and_in_chunks():
    """section 8 f2: the evidence bootstrap (sampler.py:913) without the [B, n] index matrix -- (a) host rows consumed in
    chunks of 256 give the same replicates, bit for bit, as the one-shot matrix; (b) device-drawn indices give the same
    distribution (mean / std of the replicates within Monte-Carlo error) and do not depend on how the launch is split."""
    from pocomc_b200.tools import lse_device
    rng = np.random.default_rng(5)
    logw = rng.normal(size=3000) * 2.0
    n, B = len(logw), 700
    np.random.seed(3)
    full = np.stack([np.random.choice(n, n) for _ in range(B)])
    z0, b0 = lse_device(logw, boot_idx=full)
    np.random.seed(3)
    z1, b1 = lse_device(logw, n_boot=B, boot_rows=lambda k: np.stack([np.random.choice(n, n) for _ in range(k)]))
    assert z0 == z1
    np.testing.assert_array_equal(b0, b1)
    z2, b2 = lse_device(logw, n_boot=4000, seed=11)
    _, b3 = lse_device(logw, n_boot=4000, seed=11)
    np.testing.assert_array_equal(b2, b3)                      # counter-based: reproducible
    assert z2 == z0
    assert abs(np.mean(b2) - np.mean(b0)) < 4 * np.std(b0) / np.sqrt(B)
    assert abs(np.std(b2) / np.std(b0) - 1.0) < 0.15
    _, b4 = lse_device(logw, n_boot=4000, seed=12)
    assert not np.array_equal(b2, b4)
