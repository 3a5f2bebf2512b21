#!/bin/bash
mkdir -p gpurun_out
timeout 250 python tests/nsf_truth.py > gpurun_out/r2bj_nsf_truth_lean.log 2>&1; echo "rc=$?"
PMC_B200_LIBPATH=$PWD/build/alt/libpmc_b200_refhead.so timeout 250 python tests/nsf_truth.py > gpurun_out/r2bj_nsf_truth_refhead.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for f in ("lean","refhead"):
    print("==", f)
    for line in open(f"gpurun_out/r2bj_nsf_truth_{f}.log"):
        if not line.startswith("{"): print(line.strip()[:200]); continue
        r=json.loads(line)
        print(r["d"],r["flow"],"in",r["input_scale"],"w",r["weight_scale"],"inv" if r["inverse"] else "fwd", "| tri x p999 %.1e max %.1e ladj p999 %.1e max %.1e bad %d/%d | ffma x p999 %.1e max %.1e ladj p999 %.1e max %.1e bad %d/%d"%(
            r["tri"]["err_x_p999"],r["tri"]["err_x_max"],r["tri"]["err_ladj_p999"],r["tri"]["err_ladj_max"],r["tri"]["rows_x_gt_5e4"],r["tri"]["rows_ladj_gt_5e3"],
            r["ffma"]["err_x_p999"],r["ffma"]["err_x_max"],r["ffma"]["err_ladj_p999"],r["ffma"]["err_ladj_max"],r["ffma"]["rows_x_gt_5e4"],r["ffma"]["rows_ladj_gt_5e3"]))
PY
