timeout 300 python -m pytest tests/test_gpu_flow.py tests/test_gpu_sampler.py -x -q 2>&1 | tail -3
timeout 300 python tests/train_bench.py 2>&1 | grep fused
timeout 300 python tests/run_profile.py rosen10 1000 2>&1 | tail -1
timeout 300 python tests/run_profile.py gauss32 10000 2>&1 | tail -1
