#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -rf --tb=short > gpurun_out/r02d_pytest.log 2>&1
QSB_WS=1 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 1200 -rf --tb=line -k "32 or funnel70 or schools62 or c5 or c4 or mixture" > gpurun_out/r02d_pytest_ws.log 2>&1
for w in c5 c4; do
  QSB_WS=1 python bench.py --workload $w --steps 20 --no-cpu-baseline --ess-steps 0 > gpurun_out/r02d_bench_${w}_ws2.log 2>&1
  QSB_WS=1 QSB_WS_NPROD=1 python bench.py --workload $w --steps 20 --no-cpu-baseline --ess-steps 0 > gpurun_out/r02d_bench_${w}_ws1.log 2>&1
done
cmd="python bench.py --workload c5 --steps 2 --warmup 3 --burnin 20 --no-cpu-baseline --e2e-steps 0 --ess-steps 0 --roofline-steps 0"
QSB_WS=1 ncu --set full --clock-control none --import-source on -k regex:advance_ws -s 12 -c 1 -f -o gpurun_out/r02d_c5_ws $cmd > gpurun_out/r02d_ncu_c5_ws.log 2>&1
timeout 500 python tools/converge.py c4 --chains 1024 --chunk 10000 --chunks 6 > gpurun_out/r02d_conv_c4.log 2>&1
timeout 500 python tools/converge.py c5 --chains 512 --chunk 5000 --chunks 6 > gpurun_out/r02d_conv_c5.log 2>&1
tail -n 40 gpurun_out/r02d_pytest.log
