#!/bin/bash
# A/B of k_call_alleles build variants on the GPU box: tools/ab_k1.sh "WARPS:SC_OPS:CAND_CAP[:extra -D flags]" ...   (the last one stays built)
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  IFS=: read W S C X <<< "$cfg"
  touch longphase-s_b200/csrc/k_call_alleles.cu
  make -C longphase-s_b200/csrc EXTRA="-DLPS_WARPS=$W -DLPS_SC_OPS=$S -DLPS_CAND_CAP=$C $X" > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --contigs-per-gpu 1 ${AB_ARGS} > /tmp/ab.json 2>/tmp/ab.err || { tail -3 /tmp/ab.err; continue; }
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open('/tmp/ab.json'))
print(sys.argv[1], "k1_ms=%.4f frac=%.3f step_ms=%.3f fold_ms=%.3f" % (d["stage_ms"]["k_call_alleles_alone"], d["roofline"]["frac"], d["ms_per_step"], d["stage_ms"]["k_fold_edges_alone"]))
PY
done
