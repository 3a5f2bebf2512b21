"""TEST INFRASTRUCTURE -- records whole-run results of the UNMODIFIED reference Sampler
(/root/reference/pocomc, with oracle/zuko standing in for the absent third-party zuko) into
tests/golden/runs.json.  Run in the build container only: python oracle/make_golden_runs.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, "/root/reference"]
import numpy as np  # noqa: E402
import pocomc as pc  # noqa: E402
from scipy.stats import norm, uniform  # noqa: E402

CASES = {
    "box3": dict(prior=("uniform", -5, 10, 3), like="offset_gauss"),
    "gauss2": dict(prior=("norm", 0, 1, 2), like="unit_gauss"),
}
RUNS = [
    ("box3", dict(precondition=False), 3),
    ("box3", dict(precondition=False, sample="rwm", resample="syst"), 3),
    ("box3", dict(sample="rwm", flow="maf3", metric="uss", periodic=[0], reflective=[1]), 3),
    ("box3", dict(flow="maf3", periodic=[0], reflective=[1]), 4),
    ("box3", dict(flow="nsf3"), 5),
    ("gauss2", dict(flow="maf3", train_config=dict(epochs=1)), 0),
]


def likelihood(name):
    if name == "offset_gauss":
        return lambda x: -0.5 * np.sum((x - 1.0) ** 2, axis=1) / 0.09
    return lambda x: -0.5 * np.sum(x ** 2, axis=-1)


def prior_of(spec):
    kind, a, b, d = spec
    return pc.Prior([(uniform if kind == "uniform" else norm)(a, b)] * d)


def main():
    out = []
    for case, kw, seed in RUNS:
        c = CASES[case]
        s = pc.Sampler(prior_of(c["prior"]), likelihood(c["like"]), vectorize=True, n_effective=256, n_active=128,
                       random_state=seed, **kw)
        s.run(n_total=512, n_evidence=0, progress=False)
        r = s.results
        out.append(dict(case=case, kwargs=kw, seed=seed, logz=float(s.evidence()[0]), iterations=int(s.t),
                        calls=int(s.calls), beta=[float(b) for b in r["beta"]], steps=[int(v) for v in r["steps"]],
                        logz_path=[float(v) for v in r["logz"]]))
        print(case, kw, seed, out[-1]["logz"], s.t, s.calls)
    with open(os.path.join(HERE, "..", "tests", "golden", "runs.json"), "w") as f:
        json.dump(dict(cases=CASES, runs=out), f, indent=1)


if __name__ == "__main__":
    main()
