#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2bc_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2bc_pytest_gpu.log
tail -6 gpurun_out/r2bc_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bc_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2bc_smoke.log
