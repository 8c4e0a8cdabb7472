#!/bin/bash
# multi-GPU measurement set: N = $1 GPUs.  Group / comm tests, torchrun benches of the BASELINE multi-GPU configurations,
# the same through qsb_group from one host thread.
set -u
N=${1:-8}; tag=${2:-r02m}; steps=${3:-20}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.log 2>&1
python -m pytest tests/test_gpu_group.py -m gpu -q --timeout 600 -rA --tb=short > gpurun_out/${tag}_pytest_group.log 2>&1
tail -n 4 gpurun_out/${tag}_pytest_group.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps $steps --warmup 3 "${@:2}"; }
run 29511 > gpurun_out/${tag}_bench_c3_n$N.log 2> gpurun_out/${tag}_bench_c3_n$N.err
run 29512 --workload c4 > gpurun_out/${tag}_bench_c4_n$N.log 2> gpurun_out/${tag}_bench_c4_n$N.err
run 29513 --workload c5 > gpurun_out/${tag}_bench_c5_n$N.log 2> gpurun_out/${tag}_bench_c5_n$N.err
for w in c3 c4 c5; do tail -c 600 gpurun_out/${tag}_bench_${w}_n$N.log; echo; tail -n 3 gpurun_out/${tag}_bench_${w}_n$N.err | cut -c1-300; done
for w in c3 c5; do python tools/group_bench.py $w --steps $steps > gpurun_out/${tag}_group_${w}_n$N.log 2>&1; tail -n 2 gpurun_out/${tag}_group_${w}_n$N.log | cut -c1-600; done
