#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -rf --tb=short > gpurun_out/r02c_pytest.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench_c3.log 2>&1
for c in 1024 2048 4096; do
  python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/r02c_bench_c3_$c.log 2>&1
done
for w in c5 c4; do
  python bench.py --workload $w --steps 20 --no-cpu-baseline --ess-steps 0 > gpurun_out/r02c_bench_${w}.log 2>&1
  QSB_NO_WS=1 python bench.py --workload $w --steps 20 --no-cpu-baseline --ess-steps 0 > gpurun_out/r02c_bench_${w}_nows.log 2>&1
done
python bench.py --workload c2 --steps 20 --no-cpu-baseline > gpurun_out/r02c_bench_c2.log 2>&1
QSB200_LIB=$PWD/quicksand_b200/libqsb200_c2const.so python bench.py --workload c2 --steps 20 --no-cpu-baseline > gpurun_out/r02c_bench_c2_const.log 2>&1
timeout 500 python tools/converge.py c4 --chains 1024 --chunk 10000 --chunks 6 > gpurun_out/r02c_conv_c4.log 2>&1
timeout 500 python tools/converge.py c5 --chains 512 --chunk 5000 --chunks 6 > gpurun_out/r02c_conv_c5.log 2>&1
tail -n 30 gpurun_out/r02c_pytest.log
