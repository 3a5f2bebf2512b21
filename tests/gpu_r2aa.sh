#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler_sharded.py -m gpu -q -x > gpurun_out/r2aa_sharded_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2aa_sharded_tests.log
grep -n "Error\|error\|FAILED\|passed\|failed" gpurun_out/r2aa_sharded_tests.log | head -20
for p2p in 1 0; do
for sc in weak strong; do
PMC_B200_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --scaling $sc --no-aux --no-cpu-baseline 2> gpurun_out/r2aa_${sc}_p2p${p2p}.err | grep '^{' > gpurun_out/r2aa_${sc}_p2p${p2p}.json
echo "bench $sc p2p=$p2p rc=$?"; cut -c1-260 gpurun_out/r2aa_${sc}_p2p${p2p}.json; grep -v Warning gpurun_out/r2aa_${sc}_p2p${p2p}.err | tail -3
done
done
