#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2y_pytest_gpu.log
tail -15 gpurun_out/r2y_pytest_gpu.log
( D=32 N=10000 timeout 300 python tests/tri_bench.py
  D=10 N=1000 timeout 300 python tests/tri_bench.py
  D=200 N=125000 ITER=2 timeout 600 python tests/tri_bench.py ) > gpurun_out/r2y_tri_bench.log 2>&1
cut -c1-400 gpurun_out/r2y_tri_bench.log
python tests/tc_stress.py 5 2>&1 | grep -v Warn | tail -12
