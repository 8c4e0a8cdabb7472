#!/bin/bash
# 16 warps per CTA with Theta in shared memory + the permuted layout (libqsb200_cw16.so) against the default build: parity subset
# on the variant, then same-box A/B
set -u
tag=${1:-r03f}
mkdir -p gpurun_out
QSB200_LIB=$PWD/quicksand_b200/libqsb200_cw16.so python -m pytest tests -m gpu -q --timeout 600 -rf --tb=short -x -k "glm_kernel_dimension or c3_full_shape_gradient or stream_k or working_regime" > gpurun_out/${tag}_pytest_cw16.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_cw16.log
for rep in 1 2; do
for v in main cw16; do
  lib=quicksand_b200/libqsb200.so; [ $v != main ] && lib=quicksand_b200/libqsb200_$v.so
  for c in 8192 1024; do
    QSB200_LIB=$PWD/$lib python bench.py --steps 20 --chains $c --no-cpu-baseline --ess-steps 0 --e2e-steps 0 > gpurun_out/${tag}_bench_c3_${c}_${v}_$rep.log 2>&1
    python - gpurun_out/${tag}_bench_c3_${c}_${v}_$rep.log <<'PY'
import sys, json
for line in open(sys.argv[1]):
    if line.startswith('{'):
        j = json.loads(line); print(sys.argv[1], "%.5g" % j["value"], "ms %.4f" % j["ms_per_step"], "kernel %.5f ms" % j["roofline"]["kernel_ms_per_launch"], "frac %.4f" % j["roofline"]["frac"])
PY
  done
done
done
