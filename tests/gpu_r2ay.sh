#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/chain_bench.py > gpurun_out/r2ay_chain.log 2>&1; echo "chain rc=$?"
head -16 gpurun_out/r2ay_chain.log
timeout 600 python -m pytest tests/test_gpu_mcmc.py tests/test_gpu_sampler.py -m gpu -q -x 2>&1 | tail -3
