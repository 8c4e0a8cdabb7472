#!/bin/bash
# A/B of the interleaved Gaussian fill (QSB_FILL_U = 1 / 2 / 4) on the fused transition kernels + parity subset + c3 variance probe
set -u
tag=${1:-r02k}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 1200 -rf --tb=short -x -k "rng_stream or gaussian_fill or full_dimension or drift_per_step or harm_per_step or mixture_harm or hmc_fixed" > gpurun_out/${tag}_pytest_subset.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_subset.log
for v in u1 u2 u4; do
  lib=quicksand_b200/libqsb200_$v.so; [ $v = u2 ] && lib=quicksand_b200/libqsb200.so
  for w in c5 c4 c2; do
    QSB200_LIB=$PWD/$lib python bench.py --workload $w --steps 20 --no-cpu-baseline --ess-steps 0 --e2e-steps 0 > gpurun_out/${tag}_bench_${w}_$v.log 2>&1
    python - gpurun_out/${tag}_bench_${w}_$v.log <<'PY'
import sys, json
for line in open(sys.argv[1]):
    if line.startswith('{'):
        j = json.loads(line); print(sys.argv[1], "%.4g" % j["value"], "ms %.4f" % j["ms_per_step"], "frac %.4f" % j["roofline"]["frac"])
PY
  done
done
python tools/c3_variance_probe.py > gpurun_out/${tag}_c3_variance_probe.log 2>&1
tail -n 6 gpurun_out/${tag}_c3_variance_probe.log
