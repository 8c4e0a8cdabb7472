#!/bin/bash
# ncu --set full capture of the chain-parallel advance kernel (one timed launch per workload).
# usage (on the GPU box): bash tools/profile_adv.sh <tag> [workloads...]
set -u
tag=${1:-r01}; shift
wls=${@:-c4 c5 c2}
mkdir -p gpurun_out
for w in $wls; do
  cmd="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0"
  $cmd > gpurun_out/plain_$w.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:advance_kernel -s 4 -c 1 -f -o gpurun_out/${tag}_${w}_adv $cmd > gpurun_out/ncu_$w.log 2>&1
  echo "$w rc=$?"
done
