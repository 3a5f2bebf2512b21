#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2au_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2au_pytest_gpu.log
tail -3 gpurun_out/r2au_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2au_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2au_smoke.log
timeout 900 python bench.py > gpurun_out/r2au_bench_cfg1.json 2> gpurun_out/r2au_bench_cfg1.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/r2au_bench_cfg1.json
