#!/bin/bash
# N-GPU scaling lines: weak (10 000 particles per GPU) with the peer-memory exchange and with NCCL, strong (640 000 particles)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for mode in "weak 1" "weak 0" "strong 1"; do
set -- $mode
PMC_B200_P2P=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --scaling $1 --no-aux --no-cpu-baseline 2> gpurun_out/r2af_${1}_p2p${2}_${N}gpu.err | grep '^{' > gpurun_out/r2af_${1}_p2p${2}_${N}gpu.json
echo "bench $1 p2p=$2 N=$N rc=$?"; cut -c1-200 gpurun_out/r2af_${1}_p2p${2}_${N}gpu.json; grep -v "Warning\|warn\|return func\|^\*\*\*\|OMP_NUM" gpurun_out/r2af_${1}_p2p${2}_${N}gpu.err | tail -4
done
