#!/bin/bash
# round-2 measurement set, part 1: new tests, benches of every workload, reference arm, c3 launch list + one full capture
set -u
tag=${1:-r02g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -rf --tb=short -k "moments or health or group or conditions" > gpurun_out/${tag}_pytest_new.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_c3.log 2>&1
for c in 1024 2048 4096; do
  python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/${tag}_bench_c3_$c.log 2>&1
done
python bench.py --steps 20 --chains 1024 --no-cpu-baseline --ess-steps 0 --e2e-steps 5 --invariant > gpurun_out/${tag}_bench_c3_1024_inv.log 2>&1
for w in c2 c4 c5; do python bench.py --workload $w --steps 20 > gpurun_out/${tag}_bench_$w.log 2>&1; done
python bench.py --impl reference --steps 20 > gpurun_out/${tag}_bench_ref.log 2>&1
cmd="python bench.py --steps 2 --warmup 3 --burnin 5 --no-cpu-baseline --e2e-steps 1 --ess-steps 0 --roofline-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_c3_launches.csv $cmd > gpurun_out/${tag}_ncu_c3_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:glm_grad_kernel -s 60 -c 1 -f -o gpurun_out/${tag}_c3_grad $cmd > gpurun_out/${tag}_ncu_c3_full.log 2>&1
cmd1k="python bench.py --steps 2 --warmup 3 --burnin 5 --no-cpu-baseline --e2e-steps 1 --ess-steps 0 --roofline-steps 1 --chains 1024"
ncu --set full --clock-control none -k regex:glm_grad_kernel -s 60 -c 1 -f -o gpurun_out/${tag}_c3_grad_1024 $cmd1k > gpurun_out/${tag}_ncu_c3_1024.log 2>&1
du -sh gpurun_out
tail -n 5 gpurun_out/${tag}_pytest_new.log
