"""CPU: pin the oracle restatement (oracle/) against vectors recorded from the unmodified
reference (tests/golden/, made by oracle/make_golden.py) and SURVEY App. C known answers."""
import numpy as np
import pytest
import torch

import smc_ref as O
import flow_ref as F
from conftest import flow_param_list

F64 = dict(rtol=1e-12, atol=1e-12)


def _scaler(g, tr):
    return O.ScalerParams(g["low"], g["high"], g[f"{tr}_mu"], g[f"{tr}_sigma"], logit=(tr == "logit"))


@pytest.mark.parametrize("tr", ["probit", "logit"])
def test_scaler_matches_reference(golden, tr):
    g = golden("scaler")
    p = O.scaler_fit(g["x"], g["low"], g["high"], logit=(tr == "logit"))
    np.testing.assert_allclose(p.mu, g[f"{tr}_mu"], **F64)
    np.testing.assert_allclose(p.sigma, g[f"{tr}_sigma"], **F64)
    p = _scaler(g, tr)
    np.testing.assert_allclose(O.scaler_forward(g["x"], p), g[f"{tr}_fwd"], **F64)
    with np.errstate(all="ignore"):
        x, ld = O.scaler_inverse(g["u_probe"], p)
        x32, ld32 = O.scaler_inverse(g["u_probe"].astype(np.float32), p)
    np.testing.assert_allclose(x, g[f"{tr}_inv_x"], **F64)
    np.testing.assert_allclose(ld, g[f"{tr}_inv_logdet"], **F64)
    np.testing.assert_allclose(x32, g[f"{tr}_inv32_x"], **F64)
    np.testing.assert_allclose(ld32, g[f"{tr}_inv32_logdet"], **F64)
    # round trip (reference tests/test_scaler.py:56-130)
    xr, _ = O.scaler_inverse(O.scaler_forward(g["x"], p), p)
    np.testing.assert_allclose(xr, g["x"], rtol=1e-9, atol=1e-9)


def test_boundary_conditions(golden):
    g = golden("scaler")
    p = _scaler(g, "probit")
    np.testing.assert_array_equal(O.apply_boundary_conditions(g["bc_in"], p, [3], [4]), g["bc_out"])


def test_scaler_kats():
    # SURVEY App. C
    p = O.ScalerParams(np.array([0, 0, -np.inf]), np.array([1, np.inf, np.inf]), np.zeros(3), np.array([1, 2, .5]))
    x, ld = O.scaler_inverse(np.array([[0., 0, 0], [1, -1, 2]]), p)
    np.testing.assert_allclose(x, [[0.5, 1, 0], [0.8413447460685429, 0.1353352832366127, 1]], rtol=1e-15)
    np.testing.assert_allclose(ld, [-0.9189385332046727, -3.4189385332046727], rtol=1e-15)
    p = O.ScalerParams(np.array([0.]), np.array([1.]), np.zeros(1), np.ones(1), logit=True)
    x, ld = O.scaler_inverse(np.array([[0.5]]), p)
    np.testing.assert_allclose(x, [[0.6224593312018546]], rtol=1e-15)
    np.testing.assert_allclose(ld, [-1.4481539683602134], rtol=1e-15)


def test_ps_logw_and_sample_sizes(golden):
    g = golden("smc")
    for tag, b, nz in (("b1", 1.0, True), ("b05", 0.5, True), ("b03raw", 0.3, False), ("b0", 0.0, True)):
        lw, lz = O.ps_logw(g["logl"], g["beta"], g["logz"], b, normalize=nz)
        np.testing.assert_array_equal(lw, g[f"logw_{tag}"])       # streaming form is bit-identical
        assert lz == g[f"logz_{tag}"]
    lw = g["logw_b05"]
    w = np.exp(lw - lw.max())
    assert O.ess(w) == g["ess_b05"]
    assert O.uss(w) == g["uss_b05"]
    assert O.uss(w, 100) == g["uss_b05_k100"]
    assert O.compute_ess(lw) == g["compute_ess_b05"]
    assert O.increment_logz(lw) == g["increment_logz_b05"]


def test_smc_kats(golden):
    g = golden("smc")
    assert g["kat_ess"] == pytest.approx(3.333333333333333, rel=1e-15)
    assert O.ess(np.array([1., 2, 3, 4])) == g["kat_ess"]
    assert O.compute_ess(np.log([1, 2, 3, 4])) == pytest.approx(0.8333333333333333, rel=1e-15)
    assert O.uss(np.ones(512), k=256) == pytest.approx(201.60809550983944, rel=1e-14)
    np.random.seed(0)
    np.testing.assert_array_equal(O.systematic_resample(4, np.array([0.6, 0.2, 0.15, 0.05]), np.random.random()),
                                  [0, 0, 1, 2])
    np.testing.assert_array_equal(g["kat_syst"], [0, 0, 1, 2])
    np.random.seed(0)
    np.testing.assert_array_equal(O.multinomial_resample([0.6, 0.2, 0.15, 0.05], np.random.random_sample(8)),
                                  [0, 1, 1, 0, 0, 1, 0, 2])
    idx, w = O.trim_weights(np.arange(10), np.arange(1., 11.), 0.9, 10)
    np.testing.assert_array_equal(idx, np.arange(2, 10))
    np.testing.assert_allclose(w, np.arange(3., 11.) / 52, rtol=1e-15)
    # particles KAT
    lw, lz = O.ps_logw(np.array([[-1., -2], [-0.5, -3]]), np.array([0, 0.5]), np.array([0, -0.7]), 1.0, False)
    np.testing.assert_allclose(lz, -1.358951201540815, rtol=1e-15)


def test_trim_and_resample(golden):
    g = golden("smc")
    lw = g["logw_b05"]
    w = np.exp(lw - lw.max())
    wn = w / w.sum()
    for tag, e, b in (("a", 0.99, 1000), ("b", 0.9, 50), ("c", 0.999, 200)):
        idx, wt = O.trim_weights(np.arange(len(wn)), wn.copy(), e, b)
        np.testing.assert_array_equal(idx, g[f"trim_{tag}_idx"])
        np.testing.assert_array_equal(wt, g[f"trim_{tag}_w"])
        keep, wt2, _ = O.trim_weights_sorted(wn.copy(), e, b)
        np.testing.assert_array_equal(np.nonzero(keep)[0], g[f"trim_{tag}_idx"])
        np.testing.assert_allclose(wt2, g[f"trim_{tag}_w"], rtol=1e-13)
    np.testing.assert_array_equal(O.systematic_resample(500, wn, float(g["syst_u0"])), g["syst_idx"])
    np.testing.assert_array_equal(O.multinomial_resample(wn, g["mult_r"]), g["mult_idx"])


def test_geometry(golden):
    g = golden("geometry")
    out = O.geometry_fit(g["theta"], g["w"], float(g["u0"]))
    for k in ("normal_mean", "normal_cov", "t_mean", "t_cov", "t_nu"):
        np.testing.assert_allclose(out[k], g[k], **F64)
    out = O.geometry_fit(g["theta"], None, None)
    for k in ("normal_mean", "normal_cov", "t_mean", "t_cov", "t_nu"):
        np.testing.assert_allclose(out[k], g["nw_" + k], **F64)


def test_reweight_control_flow(golden):
    g = golden("reweight")
    for j in range(3):
        b, lz, e, w, _ = O.reweight_select_beta(g[f"t{j}_logl"], g[f"t{j}_beta"], g[f"t{j}_logz"],
                                                int(g[f"t{j}_n_eff_in"]))
        assert b == g[f"t{j}_beta_out"]
        np.testing.assert_allclose(lz, g[f"t{j}_logz_out"], rtol=1e-13)
        np.testing.assert_allclose(e, g[f"t{j}_ess_out"], rtol=1e-12)
        idx, wt = O.trim_weights(np.arange(len(w)), w.copy(), 0.99, 1000)
        np.testing.assert_allclose(wt, g[f"t{j}_weights"], rtol=1e-12)
        np.testing.assert_array_equal(g[f"t{j}_logl"].reshape(-1)[idx], g[f"t{j}_logl_sel"])


def _oracle_flow(g, prefix, preset, d):
    return F.load_params(F.make_flow(d, preset), flow_param_list(g, prefix))


@pytest.mark.parametrize("tag", ["free", "bounded", "nsf"])
def test_mcmc_kernels_match_reference(golden, tag):
    g = golden("mcmc_" + tag)
    d = g["x"].shape[1]
    flow = F.NumpyFlow(_oracle_flow(g, "", "nsf3" if tag == "nsf" else "maf3", d))
    th0, lf0 = flow.forward(g["u"])
    np.testing.assert_array_equal(th0, g["theta0"])
    np.testing.assert_array_equal(lf0, g["ldjf0"])
    scaler = O.ScalerParams(g["low"], g["high"], g["mu"], g["sigma"])
    Ci = g["Cinv"]

    def loglike(x):
        return -0.5 * np.einsum("ki,ij,kj->k", x, Ci, x)

    from scipy.stats import norm, uniform
    dists = [uniform(-6, 12) if k else norm(0, 3) for k in g["prior_kind"]]

    def logprior(x):
        return sum(dd.logpdf(x[:, i]) for i, dd in enumerate(dists))

    state = dict(u=g["u"], x=g["x"], logdetj=g["logdetj"], logl=g["logl"], logp=g["logp"], beta=float(g["beta"]))
    for key in ("tpcn_flow_nufit", "tpcn_flow_nu5", "rwm_flow_nufit", "tpcn_nufit", "tpcn_nu5", "rwm_nufit"):
        if f"{key}_out_steps" not in g:
            assert tag == "nsf" and not key.rsplit("_", 1)[0].endswith("flow")      # flow-free kernels: recorded once
            continue
        kind = key.rsplit("_", 1)[0]
        steps = int(g[f"{key}_out_steps"])
        tp = kind.startswith("tpcn")
        noise = O.ReplayNoise([O.Noise(g[f"{key}_g"][i] if tp else None, g[f"{key}_z"][i], g[f"{key}_r"][i])
                               for i in range(steps)])
        geo = dict(t_mean=g[f"{key}_t_mean"], t_cov=g[f"{key}_t_cov"], t_nu=float(g[f"{key}_t_nu"]),
                   normal_cov=g[f"{key}_normal_cov"])
        res = O.mcmc_kernel(kind, state, loglike, logprior, scaler, geo,
                            dict(n_max=6, n_steps=3, proposal_scale=2.38 / d ** 0.5), flow=flow, noise=noise)
        assert res["steps"] == steps, key
        assert res["calls"] == int(g[f"{key}_out_calls"]), key
        for k in ("u", "x", "logdetj", "logl", "logp"):
            np.testing.assert_allclose(res[k], g[f"{key}_out_{k}"], rtol=1e-9, atol=1e-9, err_msg=f"{key}:{k}")
        for k in ("efficiency", "accept", "proposal_scale"):
            np.testing.assert_allclose(res[k], g[f"{key}_out_{k}"], rtol=1e-12, err_msg=f"{key}:{k}")
        # the global-stream noise source reproduces the recorded draws (SURVEY H3 identities)
        np.random.seed(int(g[f"{key}_seed"]))
        gg, zz = O.GlobalNumpyNoise().gamma_normal(len(g["x"]), d, (d + geo["t_nu"]) / 2 if tp else None)
        if tp:
            np.testing.assert_array_equal(gg, g[f"{key}_g"][0])
        np.testing.assert_array_equal(zz, g[f"{key}_z"][0])


@pytest.mark.parametrize("preset,d", [("maf3", 4), ("nsf3", 5), ("maf6", 10)])
def test_flow_wrapper_matches_reference(golden, preset, d):
    g = golden("flow")
    pre = preset + "_"
    flow = _oracle_flow(g, pre, preset, d)
    x = torch.tensor(g[pre + "x"])
    with torch.no_grad():
        nf = flow()
        z, ladj = nf.transform.call_and_ladj(x)
        xi, li = nf.transform.inv.call_and_ladj(x)
        lp = nf.log_prob(x)
    np.testing.assert_array_equal(z.numpy(), g[pre + "z"])
    np.testing.assert_array_equal(ladj.numpy(), g[pre + "ladj"])
    np.testing.assert_array_equal(xi.numpy(), g[pre + "inv_x"])
    np.testing.assert_array_equal(li.numpy(), g[pre + "inv_ladj"])
    np.testing.assert_array_equal(lp.numpy(), g[pre + "logprob"])
    # properties the reference's own tests pin (tests/test_flow.py:75-88,153-166)
    with torch.no_grad():
        xr, lr = nf.transform.inv.call_and_ladj(z)
    tol = 1e-5 if preset.startswith("maf") else 1e-4
    assert torch.allclose(xr, x, atol=tol)
    assert torch.allclose(lr, -ladj, atol=tol * 10)


@pytest.mark.parametrize("tag", ["w", "nw"])
def test_flow_fit_matches_reference(golden, tag):
    g = golden("flow_fit")
    torch.set_num_threads(1)
    flow = F.load_params(F.make_flow(4, "maf3"), flow_param_list(g, f"{tag}_init_"))
    w = torch.tensor(g["w"]) if tag == "w" else None
    torch.manual_seed(14)
    hist = F.fit(flow, torch.tensor(g["data"]), weights=w, validation_split=0.5, epochs=4, batch_size=64,
                 patience=4, shuffle=True, clip_grad_norm=1.0)
    np.testing.assert_allclose(hist["loss"], g[f"{tag}_loss"], rtol=1e-6)
    np.testing.assert_allclose(hist["val_loss"], g[f"{tag}_val_loss"], rtol=1e-6)
    for p, ref in zip(flow.parameters(), flow_param_list(g, f"{tag}_final_")):
        np.testing.assert_allclose(p.detach().numpy(), ref, rtol=1e-5, atol=1e-6)
