#!/bin/bash
# A/B of the fold kernels (16 vs 32 lanes per node) on the weak one-contig bench, then the GPU tests on the default
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
ONE="--workload weak --contigs-per-gpu 1 --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --no-e2e"
for mode in wide narrow; do
  if [ $mode = wide ]; then export LPS_FOLD_WIDE=1; else unset LPS_FOLD_WIDE; fi
  timeout 300 python bench.py $ONE > $O/fold_$mode.json 2> $O/fold_$mode.err || { echo "bench failed $mode"; tail -3 $O/fold_$mode.err; continue; }
  python - $mode $O/fold_$mode.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "fold_ms=%.4f k1_ms=%.4f step_ms=%.3f" % (d["stage_ms"]["k_fold_edges_alone"], d["stage_ms"]["k_call_alleles_alone"], d["ms_per_step"]))
PY
done
unset LPS_FOLD_WIDE
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_m.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest_m.log
