#!/bin/bash
# ncu of the c3 gradient kernel (current code): launch list + one full capture with sources, pages exported as CSV
set -u
tag=${1:-r03a}
mkdir -p gpurun_out /tmp/ncu
cmd="python bench.py --steps 2 --warmup 3 --burnin 5 --no-cpu-baseline --e2e-steps 1 --ess-steps 0 --roofline-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_c3_launches.csv $cmd > gpurun_out/${tag}_ncu_c3_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:glm_grad_kernel -s 60 -c 1 -f -o gpurun_out/${tag}_c3_grad $cmd > gpurun_out/${tag}_ncu_c3_full.log 2>&1
ncu -i gpurun_out/${tag}_c3_grad.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${tag}_c3_grad_source.csv.gz
ls -la gpurun_out
