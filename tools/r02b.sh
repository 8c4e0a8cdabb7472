#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -rf --tb=short > gpurun_out/r02b_pytest.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_c3.log 2>&1
QSB_NO_PDL=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --ess-steps 0 > gpurun_out/r02b_bench_c3_nopdl.log 2>&1
for c in 1024 2048 4096; do
  python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/r02b_bench_c3_$c.log 2>&1
done
QSB_NO_PDL=1 python bench.py --steps 20 --chains 1024 --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/r02b_bench_c3_1024_nopdl.log 2>&1
python bench.py --steps 20 --chains 1024 --no-cpu-baseline --ess-steps 0 --e2e-steps 5 --invariant > gpurun_out/r02b_bench_c3_1024_inv.log 2>&1
python bench.py --workload c2 --steps 20 --no-cpu-baseline > gpurun_out/r02b_bench_c2.log 2>&1
QSB_NO_G4=1 python bench.py --workload c2 --steps 20 --no-cpu-baseline > gpurun_out/r02b_bench_c2_g1.log 2>&1
timeout 400 python tools/converge.py c4 --chains 1024 --chunk 10000 --chunks 6 > gpurun_out/r02b_conv_c4.log 2>&1
timeout 400 python tools/converge.py c5 --chains 512 --chunk 5000 --chunks 6 > gpurun_out/r02b_conv_c5.log 2>&1
tail -n 30 gpurun_out/r02b_pytest.log
