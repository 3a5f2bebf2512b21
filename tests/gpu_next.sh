#!/bin/bash
# First GPU call of the next round: validate and time the register-blocked bulk/tip
# sweep variants, and capture what the wavefront model needs.  Run on the GPU box from the repo root:
#   gpurun --timeout 600 -- 'bash tests/gpu_next.sh'
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
PMC_B200_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_zz_gpu_experimental.py -x -q --timeout 60 2>&1 | tail -4 | tee gpurun_out/next_experimental_tests.log
PMC_B200_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_sampler_sharded.py -x -q --timeout 120 -k history 2>&1 | tail -4 | tee gpurun_out/next_sharded_history.log
{
  for n in 10000 5120 2560; do
    N=$n timeout 60 python tests/sweep_bench.py 2>&1 | grep "inverse=True" | sed "s/^/stream        /"
    for ppl in 1 2 4; do
      N=$n PMC_B200_SWEEP=tip PMC_TIP_PPL=$ppl timeout 60 python tests/sweep_bench.py 2>&1 | grep "inverse=True" | sed "s/^/tip PPL=$ppl     /"
    done
  done
} | tee gpurun_out/next_sweep_times.log
# training step: default tiling vs split-K hidden GEMMs
{ timeout 100 python tests/train_bench.py 2>&1 | grep fused | sed 's/^/default  /'; PMC_TRAIN_SPLITK=1 timeout 100 python tests/train_bench.py 2>&1 | grep fused | sed 's/^/split-K  /'; } | tee gpurun_out/next_train_times.log
# one full capture each: default stream kernel, bulk/tip at 2 particles per lane
timeout 120 ncu --set full --clock-control none --import-source on -k regex:made_sweep_stream -c 1 -f -o gpurun_out/next_sweep_stream python tests/sweep_bench.py > /dev/null 2>&1
PMC_B200_SWEEP=tip PMC_TIP_PPL=2 timeout 120 ncu --set full --clock-control none --import-source on -k regex:made_sweep_tip -c 1 -f -o gpurun_out/next_sweep_tip2 python tests/sweep_bench.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
