"""bench.py's host-side contract (no GPU): the reference arm's JSON line and the roofline helpers."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: the unmodified reference's mcmc.py (oracle/_ref when staged, else /root/reference or the
    oracle port) on the host cores; one JSON line with impl / cpu_baseline / e2e (no copies) and the arm's own config."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "0", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "particle-steps/s" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "10-D Rosenbrock" in d["config"]["workload"] and d["config"]["flow"] == "maf6"


def test_roofline_helpers():
    import bench
    from pocomc_b200 import made_layout as ML, tri_layout as TL
    maf = TL.build_tri(32, 128, 3, 6, ML.KIND_AFFINE)
    nsf = TL.build_tri(32, 128, 3, 6, ML.KIND_RQS)
    f_maf, f_nsf = bench.tri_issued_flop(maf.meta, 10000), bench.tri_issued_flop(nsf.meta, 10000)
    # 79 tiles x 6 transforms x 3 passes of dense blocks: ~1.05 MFLOP per particle for maf6; the spline
    # flow issues more (23 output rows per feature instead of 2, and four windows to initialise)
    assert 0.9e6 < f_maf / (79 * 128) < 1.3e6 and f_nsf > 2 * f_maf
    assert bench.tri_issued_flop(maf.meta, 128) * 79 == f_maf
    old = bench.FLOW
    try:
        bench.FLOW = "maf6"
        t, src = bench.committed_traffic("made_sweep_tri_kernel<true> (flow inverse)", 10000, 32)
        assert t and "r2w_tri_d32" in src
        bench.FLOW = "nsf6"
        t2, src2 = bench.committed_traffic("made_sweep_tri_kernel<true, rqs>", 10000, 32)
        assert t2 and t2 != t and "nsf6" in src2
        t3, note = bench.committed_traffic("made_sweep_tri_kernel<true, rqs>", 12345, 32)
        assert t3 is None and "no committed ncu capture" in note
    finally:
        bench.FLOW = old
    assert os.path.exists(os.path.join(ROOT, json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[0]["source"].split(":")[0]))
