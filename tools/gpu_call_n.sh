#!/bin/bash
# sweep segment length A/B (-DLPS_SW_SEG), timing of the sweep stage and of the whole one-contig step; last value stays built; then GPU tests
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
ONE="--workload weak --contigs-per-gpu 1 --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --no-e2e"
for seg in "$@"; do
  touch longphase-s_b200/csrc/k_sweep.cu
  make -C longphase-s_b200/csrc EXTRA="-DLPS_SW_SEG=$seg" > /dev/null 2>&1 || { echo "build failed $seg"; continue; }
  timeout 300 python bench.py $ONE > $O/seg_$seg.json 2> $O/seg_$seg.err || { echo "bench failed $seg"; tail -3 $O/seg_$seg.err; continue; }
  python - $seg $O/seg_$seg.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print("seg", sys.argv[1], "sweep_ms=%.4f fold_ms=%.4f k1_ms=%.4f step_ms=%.3f fallbacks=%s" % (d["stage_ms"]["sweep"], d["stage_ms"]["k_fold_edges_alone"], d["stage_ms"]["k_call_alleles_alone"], d["ms_per_step"], d["stage_ms"]["sweep_fallbacks"]))
PY
done
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_n.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest_n.log
