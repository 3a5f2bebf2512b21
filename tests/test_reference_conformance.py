"""Drop-in conformance against the UNMODIFIED reference staged under oracle/_ref (oracle/make_ref.sh).

1. The reference's own 31 unittest cases (tests/test_{flow,scaler,prior,tools,sampler,state}.py) run with the name
   ``pocomc`` bound to ``pocomc_b200`` -- SURVEY section 4 / section 7 step 1.
2. The reference's own ``pocomc.sampler.Sampler`` (its control flow, its Particles, its scaler, its tools) runs with
   ONLY the hot path swapped for this repo's: ``Flow`` (flow.py seam) and the four MCMC kernels (mcmc.py seam), and
   must walk the same temperature ladder as the untouched reference for the same ``random_state`` -- north_star:
   "drops into sampler.py unchanged".

Both need a GPU (pocomc_b200 has no CPU compute path) and run in a subprocess so that the aliasing of ``pocomc``
cannot leak into the rest of the suite.
"""
import json
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _need_ref():
    if not os.path.isdir(os.path.join(REF, "pocomc")):
        pytest.fail("oracle/_ref is missing: run `bash oracle/make_ref.sh` in the build container (it ships with the gpurun "
                    "snapshot; /root/reference does not exist on the GPU box)")


def _run(code, cwd, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True, text=True,
                       timeout=timeout)
    return r


def test_reference_unittests_pass_on_pocomc_b200(tmp_path):
    _need_ref()
    code = f"""
        import sys, unittest
        import pocomc_b200
        from pocomc_b200 import flow, sampler, prior, scaler, tools, mcmc, particles, geometry
        sys.modules["pocomc"] = pocomc_b200
        for name, mod in dict(flow=flow, sampler=sampler, prior=prior, scaler=scaler, tools=tools, mcmc=mcmc,
                              particles=particles, geometry=geometry).items():
            sys.modules["pocomc." + name] = mod
        suite = unittest.defaultTestLoader.discover({os.path.join(REF, "tests")!r})
        res = unittest.TextTestRunner(verbosity=1).run(suite)
        print("RESULT", res.testsRun, len(res.failures), len(res.errors), len(res.skipped))
        for _, tb in res.failures + res.errors:
            print(tb)
    """
    r = _run(code, str(tmp_path))
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    assert line, r.stdout[-3000:] + r.stderr[-3000:]
    ran, fails, errors, skipped = map(int, line[0].split()[1:])
    assert (ran, fails, errors, skipped) == (31, 0, 0, 0), r.stdout[-4000:] + r.stderr[-4000:]


_DROPIN = """
    import sys, json
    sys.path.insert(0, {oracle!r})          # oracle/zuko: the reference's only missing dependency
    sys.path.insert(0, {ref!r})             # the unmodified reference package
    import numpy as np
    from scipy.stats import norm, uniform
    import pocomc                            # the REFERENCE
    assert pocomc.__file__.startswith({ref!r}), pocomc.__file__
    patched = {patched}
    if patched:
        import pocomc_b200
        from pocomc_b200 import mcmc as K
        import pocomc.sampler as S
        S.Flow = pocomc_b200.Flow            # flow.py seam
        S.preconditioned_pcn, S.preconditioned_rwm, S.pcn, S.rwm = K.preconditioned_pcn, K.preconditioned_rwm, K.pcn, K.rwm

    def loglike(x):
        return -np.sum(10.0 * (x[:, ::2] ** 2.0 - x[:, 1::2]) ** 2.0 + (x[:, ::2] - 1.0) ** 2.0, axis=1)

    out = {{}}
    for name, prior, kw in [
        ("rosen4_maf3", pocomc.Prior(4 * [uniform(-5.0, 10.0)]), dict(flow="maf3", n_active=128, n_effective=256)),
        ("gauss3_maf6_rwm", pocomc.Prior(3 * [norm(0.0, 3.0)]), dict(flow="maf6", n_active=96, n_effective=192, sample="rwm")),
    ]:
        s = pocomc.Sampler(prior, loglike, vectorize=True, random_state=3, train_config=dict(epochs=40), **kw)
        s.run(n_total=512, n_evidence=512, progress=False)
        r = s.results
        out[name] = dict(beta=[float(b) for b in r["beta"]], logz=float(s.evidence()[0]), steps=[int(v) for v in r["steps"]],
                         calls=int(s.calls), flow_cls=type(s.flow).__module__)
    print("JSON" + json.dumps(out))
"""


def test_reference_sampler_runs_over_pocomc_b200_hot_path(tmp_path):
    """The reference's Sampler with this repo's Flow + MCMC kernels patched in completes, uses them, lands on the same
    logZ as the untouched reference within Monte-Carlo error, and starts along the same temperature ladder (the first
    levels are decided before any flow arithmetic can flip an accept)."""
    _need_ref()
    runs = {}
    for patched in (False, True):
        r = _run(_DROPIN.format(oracle=os.path.join(ROOT, "oracle"), ref=REF, patched=patched), str(tmp_path), timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("JSON")]
        assert line, r.stdout[-3000:] + r.stderr[-3000:]
        runs[patched] = json.loads(line[0][4:])
    for name in runs[True]:
        a, b = runs[False][name], runs[True][name]
        assert b["flow_cls"] == "pocomc_b200.flow" and a["flow_cls"] == "pocomc.flow"
        assert b["beta"][-1] == 1.0 and a["beta"][-1] == 1.0
        # warm-up iterations (beta = 0) and the first tempered level do not depend on the flow at all
        k = sum(1 for v in a["beta"] if v == 0.0) + 1
        assert a["beta"][:k] == b["beta"][:k], (name, a["beta"][:k + 1], b["beta"][:k + 1])
        assert abs(a["logz"] - b["logz"]) < 0.5, (name, a["logz"], b["logz"])
        same_path = a["beta"] == b["beta"] and a["steps"] == b["steps"]
        print(name, "same_path", same_path, "logZ", a["logz"], b["logz"])
        if same_path:
            assert abs(a["logz"] - b["logz"]) <= 1e-5 * abs(a["logz"]), (name, a["logz"], b["logz"])
