#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_flow.py -m gpu -q -x -k "block_triangular" 2>&1 | tail -2
for b in 1 2 3 4; do
  echo "== split batch $b"
  for cfg in "32 10000 nsf6" "100 50000 nsf6" "50 50000 maf6" "200 125000 maf6"; do
    set -- $cfg
    PMC_TRI_SPLIT_BATCH=$b D=$1 N=$2 FLOW=$3 ITER=5 timeout 200 python tests/tri_bench.py 2>&1 | grep '"inverse": true' | python -c "import sys,json; [print(' ', r['flow'], r['d'], r['n'], 'tri %.1f us'%r['tri_us_p3'], 'maxdiff %.2e'%r['maxdiff_p3']) for r in map(json.loads, sys.stdin)]"
  done
done 2>&1 | tee gpurun_out/r2bf_split_batch.log
