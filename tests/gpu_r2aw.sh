#!/bin/bash
mkdir -p gpurun_out
for c in 1 2; do
CFG=$c timeout 300 python tests/e2e_timeline.py > gpurun_out/r2aw_e2e_timeline_cfg$c.log 2>&1; echo "rc=$?"; cat gpurun_out/r2aw_e2e_timeline_cfg$c.log | tail -20
done
CFG=2 CHUNKS=4 timeout 300 python tests/e2e_timeline.py > gpurun_out/r2aw_e2e_timeline_cfg2_c4.log 2>&1; echo "rc=$?"; cat gpurun_out/r2aw_e2e_timeline_cfg2_c4.log | tail -20
