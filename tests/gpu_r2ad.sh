#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/run_profile.py rosen10 1000 > gpurun_out/r2ad_run_rosen10.json 2> gpurun_out/r2ad_run_rosen10.err; echo rc=$?; cut -c1-600 gpurun_out/r2ad_run_rosen10.json
timeout 600 python tests/run_profile.py gauss32 10000 > gpurun_out/r2ad_run_gauss32.json 2> gpurun_out/r2ad_run_gauss32.err; echo rc=$?; cut -c1-600 gpurun_out/r2ad_run_gauss32.json
timeout 1500 python tests/run_profile.py mix50 50000 > gpurun_out/r2ad_run_mix50.json 2> gpurun_out/r2ad_run_mix50.err; echo rc=$?; cut -c1-600 gpurun_out/r2ad_run_mix50.json; tail -3 gpurun_out/r2ad_run_mix50.err
python tests/quickstart_pin.py gpu 0 1 2 3 > gpurun_out/r2ad_quickstart_gpu.log 2>&1; cut -c1-200 gpurun_out/r2ad_quickstart_gpu.log
