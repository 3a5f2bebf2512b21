#!/bin/bash
# final scaling lines on N GPUs: weak (10 000 particles per GPU), strong (606 208 particles in total), both with the peer-memory exchange
N=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $N ] && continue
  for sc in weak strong; do
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --scaling $sc --no-aux --no-cpu-baseline 2> gpurun_out/r2ar_${sc}_${n}gpu.err | grep '^{' > gpurun_out/r2ar_${sc}_${n}gpu.json
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --scaling $sc --no-aux --no-cpu-baseline 2> gpurun_out/r2ar_${sc}_${n}gpu.err | grep '^{' > gpurun_out/r2ar_${sc}_${n}gpu.json
    fi
    python -c "
import json
d=json.load(open('gpurun_out/r2ar_${sc}_${n}gpu.json')); print('$sc', $n, 'value %.1fM e2e %.1fM ms/step %.2f n/gpu %d' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['config']['n_particles_per_gpu']))"
  done
done
