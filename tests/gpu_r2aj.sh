#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flow.py -k "block_triangular or config_shapes or sweep_vs_oracle" -x -q > gpurun_out/r2aj_tri_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2aj_tri_tests.log
tail -4 gpurun_out/r2aj_tri_tests.log
for ex in 9 0; do
echo "== extra<=$ex"
( PMC_TRI_EXTRA=$ex D=200 N=125000 ITER=3 timeout 600 python tests/tri_bench.py; PMC_TRI_EXTRA=$ex D=100 N=50000 ITER=5 timeout 300 python tests/tri_bench.py; PMC_TRI_EXTRA=$ex D=50 N=50000 ITER=5 timeout 300 python tests/tri_bench.py; PMC_TRI_EXTRA=$ex D=32 N=10000 timeout 300 python tests/tri_bench.py ) 2>&1 | grep '"inverse": true' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['d'], d['n'], round(d['tri_us_p3']), 'us', d['maxdiff_p3'])"
done
