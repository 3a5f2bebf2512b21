timeout 300 python -m pytest tests/test_gpu_flow.py -x -q -k "blocked" 2>&1 | tail -5
for v in block; do PMC_B200_SWEEP=$v timeout 100 python tests/sweep_bench.py 2>&1 | tail -2; done
PMC_B200_SWEEP=block N=100000 timeout 100 python tests/sweep_bench.py 2>&1 | tail -2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:made_sweep_block -s 2 -c 1 -o gpurun_out/s4_block2 -f python tests/block_profile.py > gpurun_out/s4_ncu2.log 2>&1
