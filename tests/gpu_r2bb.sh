#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_flow.py -m gpu -q -x -k "block_triangular and nsf" > gpurun_out/r2bb_nsf_tri_tests.log 2>&1
echo "nsf tri tests rc=$?"; tail -25 gpurun_out/r2bb_nsf_tri_tests.log | cut -c1-220
timeout 200 python tests/nsf_bench.py 2>&1 | tee gpurun_out/r2bb_nsf_bench.log
