#!/bin/bash
# what do the QUAL gathers cost in the end-to-end leg?  A build without them (wrong results, gate off) against the normal one; the normal build stays
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in noqual normal; do
  touch longphase-s_b200/csrc/k_call_alleles.cu
  if [ $v = noqual ]; then EX="-DLPS_DEBUG_NO_QUAL"; export LPS_BENCH_DEBUG_NO_GATE=1; else EX=""; unset LPS_BENCH_DEBUG_NO_GATE; fi
  make -C longphase-s_b200/csrc EXTRA="$EX" > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  timeout 600 python bench.py --no-cpu-baseline --no-other-paths > $O/e2e_$v.json 2> $O/e2e_$v.err || { echo "bench failed $v"; tail -3 $O/e2e_$v.err; continue; }
  python - $v $O/e2e_$v.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "step_ms=%.3f e2e_ms=%.2f h2d=%d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"]), d.get("invalid"))
PY
done
