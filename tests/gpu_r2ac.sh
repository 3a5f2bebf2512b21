#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_flow.py -m gpu -q -x -k "layerwise or default_spline or fused_training or graph_fit or flow_fit" > gpurun_out/r2ac_train_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2ac_train_tests.log
tail -40 gpurun_out/r2ac_train_tests.log | cut -c1-250
timeout 900 python tests/train_lw_bench.py > gpurun_out/r2ac_train_lw_bench.log 2>&1
cat gpurun_out/r2ac_train_lw_bench.log | cut -c1-400
