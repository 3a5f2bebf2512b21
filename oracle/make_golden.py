"""TEST INFRASTRUCTURE -- regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box never has it):

    python oracle/make_golden.py

It puts ``oracle/`` (for the ``zuko`` restatement) and ``/root/reference`` on sys.path, imports
the reference's own modules (pocomc.mcmc / particles / scaler / tools / geometry / student /
sampler / flow) and records, per hot-path row of SURVEY.md section 8(a): the inputs, the random
draws consumed (legacy global np.random stream, re-drawn from the same seed) and the outputs.
The fixtures are committed; tests never import /root/reference.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, "/root/reference"]

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.stats import norm, uniform  # noqa: E402

import pocomc  # noqa: E402  (the reference)
from pocomc import mcmc as rmcmc  # noqa: E402
from pocomc.geometry import Geometry  # noqa: E402
from pocomc.particles import Particles  # noqa: E402
from pocomc.scaler import Reparameterize  # noqa: E402
from pocomc.student import fit_mvstud  # noqa: E402
from pocomc import tools as rtools  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(1)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def flow_state(flow):
    """Flatten a reference Flow's parameters in module order: name -> f32 array."""
    return {"p%03d" % i: p.detach().numpy().copy() for i, p in enumerate(flow.flow.parameters())}


# ------------------------------------------------------------------ scaler (a11, a12)
def golden_scaler():
    rng = np.random.default_rng(11)
    D = 6
    low = np.array([-np.inf, 0.0, -np.inf, -2.0, 1.0, -np.inf])
    high = np.array([np.inf, np.inf, 3.0, 5.0, 1.5, np.inf])
    x = np.empty((400, D))
    x[:, 0] = rng.normal(size=400) * 3
    x[:, 1] = rng.exponential(size=400) * 2
    x[:, 2] = 3.0 - rng.exponential(size=400)
    x[:, 3] = rng.uniform(-2, 5, size=400)
    x[:, 4] = rng.uniform(1.0, 1.5, size=400)
    x[:, 5] = rng.normal(size=400)
    u_probe = rng.normal(size=(300, D)) * 2.0
    u_probe[0] = 0.0
    u_probe[1] = 40.0          # saturates probit -> x hits the bound, logdet still finite
    u_probe[2] = -40.0
    u_probe[3] = 800.0         # exp overflow on the left/right-bounded dims -> inf
    out = dict(low=low, high=high, x=x, u_probe=u_probe)
    for tr in ("probit", "logit"):
        s = Reparameterize(D, bounds=np.stack([low, high], 1), transform=tr)
        s.fit(x)
        with np.errstate(all="ignore"):
            xi, ld = s.inverse(u_probe)
            xi32, ld32 = s.inverse(u_probe.astype(np.float32))
        out.update({f"{tr}_mu": s.mu, f"{tr}_sigma": s.sigma, f"{tr}_fwd": s.forward(x),
                    f"{tr}_inv_x": xi, f"{tr}_inv_logdet": ld,
                    f"{tr}_inv32_x": xi32, f"{tr}_inv32_logdet": ld32})
    # boundary conditions (scaler.py:84-157)
    s = Reparameterize(D, bounds=np.stack([low, high], 1), periodic=[3], reflective=[4])
    xb = x.copy()
    xb[:, 3] += rng.normal(size=400) * 9
    xb[:, 4] += rng.normal(size=400) * 0.8
    out["bc_in"] = xb
    out["bc_out"] = s.apply_boundary_conditions_x(xb)
    save("scaler", **out)


# ------------------------------------------------------------------ tools / particles (a14-a18)
def golden_smc():
    rng = np.random.default_rng(5)
    T, N = 12, 300
    beta = np.concatenate([[0.0, 0.0], np.sort(rng.uniform(0, 1, T - 3)), [1.0]])
    logz = np.concatenate([[0.0, 0.0], -np.cumsum(rng.uniform(0.1, 2.0, T - 2))])
    logl = -rng.chisquare(8, size=(T, N)) * 3 - 5
    p = Particles(N, 3)
    for t in range(T):
        p.update(dict(logl=logl[t], beta=beta[t], logz=logz[t]))
    out = dict(logl=logl, beta=beta, logz=logz)
    for tag, b, nz in (("b1", 1.0, True), ("b05", 0.5, True), ("b03raw", 0.3, False), ("b0", 0.0, True)):
        lw, lz = p.compute_logw_and_logz(b, normalize=nz)
        out[f"logw_{tag}"], out[f"logz_{tag}"] = lw, lz
    lw = out["logw_b05"]
    w = np.exp(lw - lw.max())
    out["ess_b05"] = rtools.effective_sample_size(w.copy())
    out["uss_b05"] = rtools.unique_sample_size(w.copy())
    out["uss_b05_k100"] = rtools.unique_sample_size(w.copy(), k=100)
    out["compute_ess_b05"] = rtools.compute_ess(lw)
    out["increment_logz_b05"] = rtools.increment_logz(lw)
    wn = w / w.sum()
    for tag, e, b in (("a", 0.99, 1000), ("b", 0.9, 50), ("c", 0.999, 200)):
        idx, wt = rtools.trim_weights(np.arange(len(wn)), wn.copy(), ess=e, bins=b)
        out[f"trim_{tag}_idx"], out[f"trim_{tag}_w"] = idx, wt
    # resampling: systematic (tools.py:136-186) and multinomial (sampler.py:702-703)
    np.random.seed(123)
    u0 = np.random.random()
    np.random.seed(123)
    out["syst_u0"] = u0
    out["syst_idx"] = rtools.systematic_resample(500, weights=wn.copy())
    np.random.seed(321)
    r = np.random.random_sample(700)
    np.random.seed(321)
    out["mult_r"] = r
    out["mult_idx"] = np.random.choice(np.arange(len(wn)), size=700, replace=True, p=wn)
    # SURVEY App. C known answers straight from the reference functions
    out["kat_ess"] = rtools.effective_sample_size(np.array([1., 2, 3, 4]))
    out["kat_compute_ess"] = rtools.compute_ess(np.log([1, 2, 3, 4]))
    out["kat_uss"] = rtools.unique_sample_size(np.ones(512), k=256)
    np.random.seed(0)
    out["kat_syst"] = rtools.systematic_resample(4, np.array([0.6, 0.2, 0.15, 0.05]))
    np.random.seed(0)
    out["kat_choice"] = np.random.choice(np.arange(4), 8, True, p=[0.6, 0.2, 0.15, 0.05])
    save("smc", **out)


# ------------------------------------------------------------------ geometry (f1)
def golden_geometry():
    rng = np.random.default_rng(9)
    D, N = 5, 600
    L = np.tril(rng.normal(size=(D, D))) + 2 * np.eye(D)
    theta = rng.standard_t(5, size=(N, D)) @ L.T + rng.normal(size=D)
    w = rng.exponential(size=N)
    w /= w.sum()
    g = Geometry()
    np.random.seed(77)
    u0 = np.random.random()
    np.random.seed(77)
    g.fit(theta, weights=w.copy())
    g2 = Geometry()
    g2.fit(theta)
    mu, S, nu = fit_mvstud(theta)
    save("geometry", theta=theta, w=w, u0=u0, normal_mean=g.normal_mean, normal_cov=g.normal_cov,
         t_mean=g.t_mean, t_cov=g.t_cov, t_nu=g.t_nu, nw_normal_mean=g2.normal_mean,
         nw_normal_cov=g2.normal_cov, nw_t_mean=g2.t_mean, nw_t_cov=g2.t_cov, nw_t_nu=g2.t_nu,
         mvstud_mu=mu, mvstud_sigma=S, mvstud_nu=nu)


# ------------------------------------------------------------------ MCMC kernels (a7-a10)
class _Recorder:
    """Wraps the likelihood so we can also store per-step proposals seen by the host callbacks."""

    def __init__(self, f):
        self.f, self.calls = f, []

    def __call__(self, x):
        self.calls.append(np.array(x, copy=True))
        return self.f(x), None


def _draw_noise(seed, n, d, shape, steps):
    """Re-draw the stream the reference consumed: per step N gammas, N*D normals, N uniforms."""
    np.random.seed(seed)
    g = np.empty((steps, n)); z = np.empty((steps, n, d)); r = np.empty((steps, n))
    for i in range(steps):
        if shape is not None:
            g[i] = np.random.standard_gamma(shape, size=n)
        z[i] = np.random.randn(n, d)
        r[i] = np.random.rand(n)
    return g, z, r


def golden_mcmc(variants=(("free", False, "maf3"), ("bounded", True, "maf3"))):
    """tag, bounded prior?, flow preset.  ("nsf", True, "nsf3") records the spline-flow variant (the reference's default
    flow family, sampler.py:169) through the two flow-preconditioned kernels -> tests/golden/mcmc_nsf.npz."""
    D, N = 4, 96
    rng = np.random.default_rng(3)
    C = 0.6 * np.ones((D, D)) + 0.4 * np.eye(D)
    Ci = np.linalg.inv(C)

    def loglike(x):
        return -0.5 * np.einsum("ki,ij,kj->k", x, Ci, x)

    for tag, bounded, preset in variants:
        if bounded:
            dists = [uniform(-6, 12), norm(0, 3), uniform(-6, 12), norm(0, 3)]
        else:
            dists = [norm(0, 3)] * D
        prior = pocomc.Prior(dists)
        np.random.seed(1)
        xs = prior.rvs(512)
        scaler = Reparameterize(D, bounds=prior.bounds)
        scaler.fit(xs)
        x0 = rng.multivariate_normal(np.zeros(D), C, size=N)
        u0 = scaler.forward(x0)
        _, ldj0 = scaler.inverse(u0)
        state = dict(u=u0, x=x0, logdetj=ldj0, logl=loglike(x0), logp=prior.logpdf(x0), beta=0.7, blobs=None)
        torch.manual_seed(4)
        flow = pocomc.Flow(D, preset)
        # a few optimiser steps so the flow is not the identity-ish initialisation
        flow.fit(torch.tensor(u0, dtype=torch.float32), epochs=8, batch_size=48)
        wrapper = rtools.flow_numpy_wrapper(flow)
        theta0, ldjf0 = wrapper.forward(u0)
        geo_t, geo_u = Geometry(), Geometry()
        np.random.seed(2)
        geo_t.fit(theta0.astype(np.float64))
        geo_u.fit(u0)
        # force a genuinely heavy-tailed nu on one variant so the gamma mixture is exercised
        out = dict(low=prior.bounds[:, 0], high=prior.bounds[:, 1], mu=scaler.mu, sigma=scaler.sigma,
                   Cinv=Ci, prior_kind=np.array([0 if isinstance(d.dist, type(norm)) else 1 for d in dists]),
                   u=u0, x=x0, logdetj=ldj0, logl=state["logl"], logp=state["logp"], beta=0.7,
                   theta0=theta0, ldjf0=ldjf0, **flow_state(flow))
        for kname, fn in (("tpcn_flow", rmcmc.preconditioned_pcn), ("rwm_flow", rmcmc.preconditioned_rwm),
                          ("tpcn", rmcmc.pcn), ("rwm", rmcmc.rwm)):
            if preset != "maf3" and not kname.endswith("_flow"):
                continue                                   # the flow-free kernels do not depend on the preset
            for nu_tag, nu in (("nufit", None), ("nu5", 5.0)):
                if nu_tag == "nu5" and not kname.startswith("tpcn"):
                    continue
                geo = geo_t if kname.endswith("_flow") else geo_u
                g = Geometry()
                g.__dict__.update(geo.__dict__)
                if nu is not None:
                    g.t_nu = nu
                rec = _Recorder(loglike)
                fd = dict(loglike=rec, logprior=prior.logpdf, scaler=scaler, flow=flow,
                          u_geometry=g, theta_geometry=g)
                od = dict(n_max=6, n_steps=3, progress_bar=None, proposal_scale=2.38 / D ** 0.5)
                seed = 100 + len(out)
                np.random.seed(seed)
                res = fn({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in state.items()}, fd, od)
                steps = int(res["steps"])
                shape = (D + g.t_nu) / 2 if kname.startswith("tpcn") else None
                gg, zz, rr = _draw_noise(seed, N, D, shape, steps)
                key = f"{kname}_{nu_tag}"
                out.update({f"{key}_seed": seed, f"{key}_g": gg, f"{key}_z": zz, f"{key}_r": rr,
                            f"{key}_t_mean": g.t_mean, f"{key}_t_cov": g.t_cov, f"{key}_t_nu": g.t_nu,
                            f"{key}_normal_cov": g.normal_cov})
                for k in ("u", "x", "logdetj", "logl", "logp"):
                    out[f"{key}_out_{k}"] = res[k]
                for k in ("efficiency", "accept", "steps", "calls", "proposal_scale"):
                    out[f"{key}_out_{k}"] = np.float64(res[k])
                out[f"{key}_xprime0"] = rec.calls[0]
        save(f"mcmc_{tag}", **out)


# ------------------------------------------------------------------ flow wrapper + training (a1-a6)
def golden_flow():
    out = {}
    for name, D in (("maf3", 4), ("nsf3", 5), ("maf6", 10)):
        torch.manual_seed(7)
        flow = pocomc.Flow(D, name)
        x = torch.randn(64, D) * 1.3
        with torch.no_grad():
            z, ladj = flow.forward(x)
            xi, ladj_i = flow.inverse(x)
            lp = flow.log_prob(x)
            torch.manual_seed(8)
            eps = torch.randn(32, D)
            torch.manual_seed(8)
            xs, lq = flow.sample(32)
        pref = f"{name}_"
        out.update({pref + "x": x.numpy(), pref + "z": z.numpy(), pref + "ladj": ladj.numpy(),
                    pref + "inv_x": xi.numpy(), pref + "inv_ladj": ladj_i.numpy(), pref + "logprob": lp.numpy(),
                    pref + "sample_eps": eps.numpy(), pref + "sample_x": xs.numpy(), pref + "sample_logq": lq.numpy()})
        out.update({pref + k: v for k, v in flow_state(flow).items()})
    save("flow", **out)

    # Flow.fit (flow.py:165-384): weighted loss, validation split, clip, early-stop bookkeeping
    D, M = 4, 384
    rng = np.random.default_rng(21)
    data = (rng.normal(size=(M, D)) @ np.array([[1, .5, 0, 0], [0, 1, .5, 0], [0, 0, 1, .5], [0, 0, 0, 1.]])).astype(np.float32)
    w = rng.exponential(size=M).astype(np.float32)
    w /= w.sum()
    fit = dict(data=data, w=w)
    for tag, weights in (("w", torch.tensor(w)), ("nw", None)):
        torch.manual_seed(13)
        flow = pocomc.Flow(D, "maf3")
        init = flow_state(flow)
        torch.manual_seed(14)
        hist = flow.fit(torch.tensor(data), weights=weights, validation_split=0.5, epochs=4, batch_size=64,
                        patience=D, annealing=False, shuffle=True, clip_grad_norm=1.0)
        fit.update({f"{tag}_init_{k}": v for k, v in init.items()})
        fit.update({f"{tag}_final_{k}": v for k, v in flow_state(flow).items()})
        fit[f"{tag}_loss"] = np.array(hist["loss"]); fit[f"{tag}_val_loss"] = np.array(hist["val_loss"])
    save("flow_fit", **fit)


# ------------------------------------------------------------------ Sampler._reweight / _resample traces (a19)
def golden_reweight():
    D = 3

    def loglike(x):
        return -0.5 * np.sum((x - 1.0) ** 2, axis=1) / 0.09

    prior = pocomc.Prior([uniform(-5, 10)] * D)
    s = pocomc.Sampler(prior, loglike, vectorize=True, n_effective=128, n_active=64, precondition=False,
                       random_state=3, dynamic=True)
    traces = []
    orig = s._reweight

    def spy(cp):
        logl = s.particles.get("logl").copy(); beta = np.array(s.particles.get("beta"), float)
        logz = np.array(s.particles.get("logz"), float)
        n_eff_in = s.n_effective
        out = orig(cp)
        traces.append(dict(logl=logl, beta=beta, logz=logz, n_eff_in=n_eff_in, n_eff_out=s.n_effective,
                           beta_out=out["beta"], logz_out=out["logz"], ess_out=out["ess"],
                           weights=out["weights"].copy(), logl_sel=out["logl"].copy()))
        return out

    s._reweight = spy
    s.run(n_total=256, n_evidence=0, progress=False)
    pick = [0, len(traces) // 2, len(traces) - 1]
    out = dict(n_active=64, dynamic_ratio=s.dynamic_ratio, final_logz=s.logz)
    for j, ti in enumerate(pick):
        for k, v in traces[ti].items():
            out[f"t{j}_{k}"] = v
    save("reweight", **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["mcmc_nsf"]:                       # added in round 2; leaves the other fixtures untouched
        golden_mcmc((("nsf", True, "nsf3"),))
        sys.exit(0)
    golden_scaler()
    golden_smc()
    golden_geometry()
    golden_mcmc()
    golden_mcmc((("nsf", True, "nsf3"),))
    golden_flow()
    golden_reweight()
