#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/quickstart_pin.py gpu 5 6 7 8 9 10 11 12 13 14 > gpurun_out/r2bk_quickstart_tri.log 2>&1; echo "tri rc=$?"
PMC_B200_TRI_RQS_MIN_H=100000 timeout 200 python tests/quickstart_pin.py gpu 5 6 7 8 9 10 11 12 13 14 > gpurun_out/r2bk_quickstart_ffma.log 2>&1; echo "ffma rc=$?"
python - <<'PY'
import json, numpy as np
for f in ("tri","ffma"):
    rs=[json.loads(l) for l in open(f"gpurun_out/r2bk_quickstart_{f}.log") if l.startswith("{")]
    z=np.array([r["logz"] for r in rs]); e=np.array([r["err"] for r in rs])
    print(f, "n", len(rs), "logz", np.round(z,2).tolist(), "err", np.round(e,2).tolist(), "mean %.3f sd %.3f  inverse-variance mean %.3f  seconds %.1f"%(z.mean(), z.std(ddof=1), (z/e**2).sum()/(1/e**2).sum(), np.mean([r["seconds"] for r in rs])))
PY
