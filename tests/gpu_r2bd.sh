#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2bd_bench_cfg1.json 2> gpurun_out/r2bd_bench_cfg1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2bd_bench_cfg1.json
timeout 600 python bench.py --flow nsf6 --no-aux > gpurun_out/r2bd_bench_cfg1_nsf6.json 2> gpurun_out/r2bd_bench_cfg1_nsf6.err; echo "bench nsf rc=$?"; cut -c1-400 gpurun_out/r2bd_bench_cfg1_nsf6.json; tail -3 gpurun_out/r2bd_bench_cfg1_nsf6.err
