#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2aq_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2aq_pytest_gpu.log
tail -4 gpurun_out/r2aq_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aq_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2aq_smoke.log
for c in 1 0 5 2 3 4; do
  timeout 900 python bench.py --config $c > gpurun_out/r2aq_bench_cfg$c.json 2> gpurun_out/r2aq_bench_cfg$c.err
  echo "bench cfg$c rc=$?"; cut -c1-120 gpurun_out/r2aq_bench_cfg$c.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2aq_bench_reference.json 2>/dev/null; cut -c1-160 gpurun_out/r2aq_bench_reference.json
