#!/bin/bash
# refresh: every bench config, ncu of the windowed tri kernel at the cfg1 / cfg4 sizes, launch list of the bench
mkdir -p gpurun_out
for c in 1 0 5 2 3 4; do
  timeout 900 python bench.py --config $c > gpurun_out/r2am_bench_cfg$c.json 2> gpurun_out/r2am_bench_cfg$c.err
  echo "bench cfg$c rc=$?"; cut -c1-200 gpurun_out/r2am_bench_cfg$c.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:made_sweep_tri -s 2 -c 1 -o gpurun_out/r2am_tri_d32 -f python tests/tri_profile.py 32 10000 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:made_sweep_tri -s 2 -c 1 -o gpurun_out/r2am_tri_d200 -f python tests/tri_profile.py 200 125000 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2am_launches.csv python bench.py --steps 1 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r2am_launches_bench.log 2>&1
ls -la gpurun_out | grep r2am
