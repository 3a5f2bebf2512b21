#!/bin/bash
for rep in 1 2; do for ex in 0 1 2; do
echo "== extra<=$ex rep $rep"
( PMC_TRI_EXTRA=$ex D=50 N=50000 ITER=10 timeout 300 python tests/tri_bench.py; PMC_TRI_EXTRA=$ex D=40 N=50000 ITER=10 FLOW=maf6 timeout 300 python tests/tri_bench.py; PMC_TRI_EXTRA=$ex D=100 N=50000 ITER=5 timeout 300 python tests/tri_bench.py ) 2>&1 | grep '"inverse": true' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['d'], d['n'], round(d['tri_us_p3']), 'us', round(d['tri_us_p1']))"
done; done
