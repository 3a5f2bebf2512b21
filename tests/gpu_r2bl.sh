#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2bl_launches.csv python bench.py --steps 1 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r2bl_launches_bench.log 2>&1
echo "ncu rc=$?"; python profiles/summarise.py launches gpurun_out/r2bl_launches.csv | tee gpurun_out/r2bl_launches_summary.txt | head -12
