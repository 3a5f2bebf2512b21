#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2ab_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2ab_pytest_gpu.log
tail -25 gpurun_out/r2ab_pytest_gpu.log | cut -c1-250
timeout 900 python tests/geometry_bench.py > gpurun_out/r2ab_geometry_bench.log 2>&1
cat gpurun_out/r2ab_geometry_bench.log | cut -c1-400
