#!/bin/bash
# 2 x B200: sharded Sampler bit-identity + weak-scaling line through the peer-memory accept kernel (staged finalize, PDL chain)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler_sharded.py -m gpu -q -x > gpurun_out/r2bg_sharded_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2bg_sharded_tests.log
grep -n "Error\|error\|FAILED\|passed\|failed" gpurun_out/r2bg_sharded_tests.log | head -10
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --scaling weak --no-aux --no-cpu-baseline 2> gpurun_out/r2bg_weak_2gpu.err | grep '^{' > gpurun_out/r2bg_weak_2gpu.json
echo "bench weak rc=$?"; cut -c1-200 gpurun_out/r2bg_weak_2gpu.json; grep -v Warning gpurun_out/r2bg_weak_2gpu.err | tail -3
