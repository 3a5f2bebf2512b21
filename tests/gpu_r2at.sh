#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flow.py -m gpu -x -q > gpurun_out/r2at_flow_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2at_flow_tests.log
tail -5 gpurun_out/r2at_flow_tests.log | cut -c1-200
timeout 300 python tests/tc_stress.py 5 2>&1 | grep -v Warn | grep "first launch\|differ\|FFMA" 
timeout 300 python tests/tc_bench.py 2>&1 | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:made_forward_tc -s 2 -c 1 -o gpurun_out/r2at_tc_forward -f python tests/tc_profile.py 1048576 > /dev/null 2>&1
