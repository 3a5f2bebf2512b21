#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcmc.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2an_mcmc_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2an_mcmc_tests.log
tail -12 gpurun_out/r2an_mcmc_tests.log | cut -c1-250
for c in 2 3 4; do
  timeout 900 python bench.py --config $c --no-cpu-baseline > gpurun_out/r2an_bench_cfg$c.json 2> gpurun_out/r2an_bench_cfg$c.err
  echo "bench cfg$c rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/r2an_bench_cfg$c.json')); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline_other'][0]['ms_per_step'], d['e2e']['value'])"
done
