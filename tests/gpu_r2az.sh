#!/bin/bash
mkdir -p gpurun_out
PMC_B200_PDL=0 timeout 300 python tests/chain_bench.py > gpurun_out/r2az_chain_pdl0.log 2>&1; echo "chain pdl0 rc=$?"
PMC_B200_PDL=1 timeout 300 python tests/chain_bench.py > gpurun_out/r2az_chain_pdl1.log 2>&1; echo "chain pdl1 rc=$?"
grep -E "rng|propose \(def|flow inv|fused|loglike|accept|device-resident|host_chunks=1|identical" gpurun_out/r2az_chain_pdl0.log
echo ---- PDL on
grep -E "rng|propose \(def|flow inv|fused|loglike|accept|device-resident|host_chunks=1|identical" gpurun_out/r2az_chain_pdl1.log
timeout 900 python -m pytest tests/test_gpu_mcmc.py tests/test_gpu_sampler.py tests/test_gpu_flow.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
