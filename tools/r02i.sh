#!/bin/bash
# source-level ncu captures of the fused transition kernels (c5, c4, c2): the per-instruction sampling page as gzipped CSV
set -u
tag=${1:-r02i}
mkdir -p gpurun_out /tmp/ncu
for w in c5 c4 c2; do
  cmdw="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 --ess-steps 0 --roofline-steps 1"
  ncu --set full --clock-control none --import-source on -k regex:advance_kernel -s 4 -c 1 -f -o /tmp/ncu/${tag}_${w}_adv $cmdw > gpurun_out/${tag}_ncu_$w.log 2>&1
  ncu -i /tmp/ncu/${tag}_${w}_adv.ncu-rep --page raw --csv > gpurun_out/${tag}_${w}_adv_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/${tag}_${w}_adv.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${tag}_${w}_adv_source.csv.gz
done
ls -la gpurun_out/ /tmp/ncu
du -sh gpurun_out
