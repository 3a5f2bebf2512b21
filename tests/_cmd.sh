timeout 100 python tests/sweep_bench.py 2>&1 | tail -2
N=100000 timeout 100 python tests/sweep_bench.py 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_flow.py -x -q -k "sweep_vs_oracle or goldens" 2>&1 | tail -2
