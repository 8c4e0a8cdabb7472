#!/bin/bash
# same-box A/B of a build variant against the main library on c3 (8 192 and 1 024 chains), alternating, twice
set -u
tag=${1:-r02u}; var=${2:-ringdiv}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -rf --tb=short -x -k "glm or c3 or logistic or stream_k or partition or smoke" > gpurun_out/${tag}_pytest_glm.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_glm.log
for rep in 1 2; do
for v in main $var; do
  lib=quicksand_b200/libqsb200.so; [ $v != main ] && lib=quicksand_b200/libqsb200_$v.so
  for c in 8192 1024; do
    QSB200_LIB=$PWD/$lib python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 5 > gpurun_out/${tag}_bench_c3_${c}_${v}_$rep.log 2>&1
    python - gpurun_out/${tag}_bench_c3_${c}_${v}_$rep.log <<'PY'
import sys, json
for line in open(sys.argv[1]):
    if line.startswith('{'):
        j = json.loads(line); print(sys.argv[1], "%.5g" % j["value"], "ms %.4f" % j["ms_per_step"], "kernel %.5f ms" % j["roofline"]["kernel_ms_per_launch"], "frac %.4f" % j["roofline"]["frac"])
PY
  done
done
done
