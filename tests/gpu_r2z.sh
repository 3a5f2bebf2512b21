#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_sampler_sharded.py -m gpu -q > gpurun_out/r2z_sharded_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2z_sharded_tests.log
grep -n "Error\|error\|FAILED\|passed\|failed" gpurun_out/r2z_sharded_tests.log | head -40
for sc in weak strong; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --scaling $sc --no-aux > gpurun_out/r2z_scale_${sc}_2gpu.json 2> gpurun_out/r2z_scale_${sc}_2gpu.err
echo "bench $sc rc=$?"; cut -c1-300 gpurun_out/r2z_scale_${sc}_2gpu.json; tail -3 gpurun_out/r2z_scale_${sc}_2gpu.err
done
