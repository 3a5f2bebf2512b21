#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,temperature.gpu,power.draw --format=csv
for rep in 1 2; do for g in 1 2; do
echo "== split_groups=$g rep=$rep"
( PMC_TRI_SPLIT_GROUPS=$g D=200 N=125000 ITER=3 timeout 600 python tests/tri_bench.py; PMC_TRI_SPLIT_GROUPS=$g D=100 N=50000 ITER=5 timeout 300 python tests/tri_bench.py; PMC_TRI_SPLIT_GROUPS=$g D=50 N=50000 ITER=5 timeout 300 python tests/tri_bench.py ) 2>&1 | grep '"inverse": true' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['d'], d['n'], round(d['tri_us_p3']), 'us')"
done; done
