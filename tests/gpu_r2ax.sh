#!/bin/bash
mkdir -p gpurun_out
for k in tpcn_propose mh_accept rng_fill scaler_inverse loglike; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r2ax_$k python tests/chain_profile.py > gpurun_out/r2ax_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/r2ax_*
