timeout 300 python -m pytest tests/test_gpu_flow.py -x -q -k "fused_training or graph_fit or fit_matches or properties" 2>&1 | tail -4
timeout 300 python tests/run_profile.py rosen10 1000 2>&1 | tail -1
timeout 300 python tests/run_profile.py gauss32 10000 2>&1 | tail -1
