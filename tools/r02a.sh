#!/bin/bash
# round 2, first GPU call: all GPU tests (incl. the new teacher-forced / group tests), the default bench, small-shard c3 runs,
# convergence exploration for c4 / c5
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02a_smi.log 2>&1
python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -60 > gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench_c3.log 2>&1
for c in 1024 2048 4096; do
  python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/r02a_bench_c3_$c.log 2>&1
  python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 5 --balanced > gpurun_out/r02a_bench_c3_${c}_bal.log 2>&1
done
timeout 300 python tools/converge.py c4 --chains 1024 --chunk 10000 --chunks 8 > gpurun_out/r02a_conv_c4.log 2>&1
timeout 300 python tools/converge.py c5 --chains 512 --chunk 5000 --chunks 8 > gpurun_out/r02a_conv_c5.log 2>&1
tail -n 5 gpurun_out/r02a_pytest.log
