#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -rf --tb=short > gpurun_out/r02e_pytest.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_c3.log 2>&1
QSB_NO_FUSED_TRAJ=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --ess-steps 0 > gpurun_out/r02e_bench_c3_nofuse.log 2>&1
for c in 1024 2048 4096; do
  python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/r02e_bench_c3_$c.log 2>&1
done
QSB_NO_FUSED_TRAJ=1 python bench.py --steps 20 --chains 1024 --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/r02e_bench_c3_1024_nofuse.log 2>&1
cmd="python bench.py --workload c5 --steps 2 --warmup 3 --burnin 20 --no-cpu-baseline --e2e-steps 0 --ess-steps 0 --roofline-steps 0"
QSB_WS=1 ncu --set full --clock-control none --import-source on -k regex:advance_ws -s 12 -c 1 -f -o gpurun_out/r02e_c5_ws $cmd > gpurun_out/r02e_ncu_c5_ws.log 2>&1
timeout 600 python tools/converge.py c4 --chains 1024 --chunk 250000 --chunks 6 --only 3 --draws 250 > gpurun_out/r02e_conv_c4_long.log 2>&1
timeout 300 python tools/converge.py c5 --chains 512 --chunk 50000 --chunks 4 --only 1 --draws 250 > gpurun_out/r02e_conv_c5_long.log 2>&1
tail -n 12 gpurun_out/r02e_pytest.log
