#!/bin/bash
run() { echo "== $*"; ( env "$@" D=200 N=125000 ITER=3 timeout 600 python tests/tri_bench.py ) 2>&1 | grep '"inverse": true' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['d'], d['n'], round(d['tri_us_p3']), 'us')"; }
run PMC_TRI_EXTRA=9
run PMC_TRI_NOMMA=1
run PMC_TRI_STAGES=2
run PMC_TRI_STAGES=2 PMC_TRI_EXTRA=0
run PMC_TRI_NOMMA=1 PMC_TRI_STAGES=2 PMC_TRI_EXTRA=0
