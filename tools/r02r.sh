#!/bin/bash
# the driver's scaling run, reproduced: bench.py (c3, strong scaling of BASELINE's 8 192 chains) at N = 1, 2, 4, 8 back to back on one box
set -u
tag=${1:-r02r}; steps=${2:-20}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_scale_n1.log 2> gpurun_out/${tag}_scale_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps $steps --warmup 3 > gpurun_out/${tag}_scale_n$N.log 2> gpurun_out/${tag}_scale_n$N.err
done
python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
base = None
for N in (1, 2, 4, 8):
    for line in open("gpurun_out/%s_scale_n%d.log" % (tag, N)):
        if line.startswith("{"):
            j = json.loads(line)
            if N == 1: base = j["value"]
            w = (j.get("secondary_weak") or {}).get("value")
            print("N=%d value %.4g (x%.3f, eff %.3f) ms %.3f e2e %.4g frac %.4f weak %s chains_total %d" % (N, j["value"], j["value"] / base, j["value"] / base / N, j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], w, j["config"]["chains_total"]))
PY
