#!/bin/bash
# Round-end measurement set on one B200: tests, smoke, benches of every workload, the reference arm, ncu launch list and
# full captures.  usage (GPU box): bash tools/final_run.sh <tag>
set -u
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_c3.log 2>&1
for w in c2 c4 c5; do python bench.py --workload $w > gpurun_out/bench_$w.log 2>&1; done
python bench.py --impl reference --steps 20 > gpurun_out/bench_ref.log 2>&1
# launch list of the default bench command (short)
cmd="python bench.py --steps 2 --warmup 3 --burnin 5 --no-cpu-baseline --e2e-steps 1"
$cmd > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_c3_launches.csv $cmd > gpurun_out/ncu_c3_list.log 2>&1
$cmd > gpurun_out/plain_c3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:glm_grad_kernel -s 60 -c 2 -f -o gpurun_out/${tag}_c3_grad $cmd > gpurun_out/ncu_c3_full.log 2>&1
[ "${SKIP_ADV:-0}" = 1 ] || bash tools/profile_adv.sh $tag c4 c5 c2
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
