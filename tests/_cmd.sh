timeout 300 python -m pytest tests/test_gpu_flow.py -x -q -k "fused_training or graph_fit or fit_matches" 2>&1 | tail -8
timeout 300 python tests/train_bench.py 2>&1 | tail -9
