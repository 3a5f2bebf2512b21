#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2ae_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2ae_pytest_gpu.log
tail -8 gpurun_out/r2ae_pytest_gpu.log | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ae_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2ae_smoke.log
timeout 900 python bench.py > gpurun_out/r2ae_bench_cfg1.json 2> gpurun_out/r2ae_bench_cfg1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2ae_bench_cfg1.json; tail -3 gpurun_out/r2ae_bench_cfg1.err
( D=32 N=10000 timeout 300 python tests/tri_bench.py ) > gpurun_out/r2ae_tri_bench.log 2>&1; cut -c1-300 gpurun_out/r2ae_tri_bench.log
