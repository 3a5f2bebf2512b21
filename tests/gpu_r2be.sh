#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:made_sweep_tri -s 2 -c 1 -f -o gpurun_out/r2be_tri_nsf6_d32 python tests/tri_profile.py 32 10000 nsf6 > gpurun_out/r2be_ncu.log 2>&1; echo "ncu rc=$?"
for v in 0 1; do PMC_TRI_NOMMA=$v D=32 N=10000 FLOW=nsf6 ITER=20 timeout 120 python tests/tri_bench.py 2>&1 | cut -c1-330; done
