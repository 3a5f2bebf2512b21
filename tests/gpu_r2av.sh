#!/bin/bash
# r2av: fused scaler + prior launch, staged finalize, chunked x' download; proposal kernel choice at D = 32
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mcmc.py tests/test_gpu_scaler.py tests/test_gpu_sampler.py tests/test_lib_abi.py -m gpu -q -x > gpurun_out/r2av_pytest_subset.log 2>&1
echo "pytest subset rc=$?"; tail -3 gpurun_out/r2av_pytest_subset.log | cut -c1-200
PMC_TPCN_TILED_MIN_D=32 timeout 300 python tests/chain_bench.py > gpurun_out/r2av_chain_tiled32.log 2>&1; echo "chain tiled32 rc=$?"
timeout 300 python tests/chain_bench.py > gpurun_out/r2av_chain_default.log 2>&1; echo "chain default rc=$?"
cat gpurun_out/r2av_chain_tiled32.log | tail -28
grep -E "propose|device-resident|e2e" gpurun_out/r2av_chain_default.log
