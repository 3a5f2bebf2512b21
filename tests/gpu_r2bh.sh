#!/bin/bash
# r2bh: final validation of the round: full GPU suite, smoke, bench lines (maf6 default, nsf6, configs 0 / 5 / 2), reference arm, quickstart pin
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2bh_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2bh_pytest_gpu.log
tail -3 gpurun_out/r2bh_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bh_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2bh_smoke.log
timeout 600 python bench.py > gpurun_out/r2bh_bench_cfg1.json 2> gpurun_out/r2bh_bench_cfg1.err; echo "bench cfg1 rc=$?"; cut -c1-150 gpurun_out/r2bh_bench_cfg1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2bh_bench_reference.json 2>/dev/null; echo "reference rc=$?"; cut -c1-160 gpurun_out/r2bh_bench_reference.json
timeout 600 python bench.py --flow nsf6 --no-aux > gpurun_out/r2bh_bench_cfg1_nsf6.json 2> gpurun_out/r2bh_bench_cfg1_nsf6.err; echo "bench nsf6 rc=$?"; cut -c1-150 gpurun_out/r2bh_bench_cfg1_nsf6.json
timeout 300 python tests/quickstart_pin.py gpu 0 1 2 3 4 > gpurun_out/r2bh_quickstart_gpu.log 2>&1; echo "quickstart rc=$?"; cut -c1-130 gpurun_out/r2bh_quickstart_gpu.log
for c in 0 5 2; do
  timeout 600 python bench.py --config $c > gpurun_out/r2bh_bench_cfg$c.json 2> gpurun_out/r2bh_bench_cfg$c.err
  echo "bench cfg$c rc=$?"; cut -c1-120 gpurun_out/r2bh_bench_cfg$c.json
done
