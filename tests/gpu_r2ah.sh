#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flow.py -k "block_triangular or config_shapes or sweep_vs_oracle" -x -q > gpurun_out/r2ah_tri_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ah_tri_tests.log
tail -4 gpurun_out/r2ah_tri_tests.log
( D=32 N=10000 timeout 300 python tests/tri_bench.py
  D=50 N=50000 ITER=5 timeout 300 python tests/tri_bench.py
  D=100 N=50000 ITER=3 timeout 300 python tests/tri_bench.py
  D=200 N=125000 ITER=2 timeout 600 python tests/tri_bench.py ) > gpurun_out/r2ah_tri_bench.log 2>&1
cut -c1-330 gpurun_out/r2ah_tri_bench.log
