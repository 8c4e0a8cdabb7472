#!/bin/bash
# full GPU test suite + benches of every workload (after the block-size balance, fetch warm-up, posterior-test comparator changes)
set -u
tag=${1:-r02n}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1200 -rf --tb=short > gpurun_out/${tag}_pytest.log 2>&1
tail -n 6 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -n 1 gpurun_out/${tag}_smoke.log
for w in c2 c4 c5; do python bench.py --workload $w --steps 20 > gpurun_out/${tag}_bench_$w.log 2>&1; done
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_c3.log 2>&1
for w in c2 c4 c5 c3; do python - gpurun_out/${tag}_bench_$w.log <<'PY'
import sys, json
for line in open(sys.argv[1]):
    if line.startswith('{'):
        j = json.loads(line); print(sys.argv[1], "%.4g" % j["value"], "ms %.4f" % j["ms_per_step"], "frac %.4f" % j["roofline"]["frac"], "e2e %.4g" % j["e2e"]["value"], "cpu %.4g" % j["cpu_baseline"]["value"])
PY
done
