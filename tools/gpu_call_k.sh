#!/bin/bash
# GPU tests + the C4 shard timing (k_window_diff) for the listed -DLPS_WD_CHUNK values (the last one stays built)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for ch in "$@"; do
  touch longphase-s_b200/csrc/k_window_diff.cu
  make -C longphase-s_b200/csrc EXTRA="-DLPS_WD_CHUNK=$ch" > /dev/null 2>&1 || { echo "build failed $ch"; continue; }
  timeout 300 python tools/som_prof.py 32 3 > $O/wd_k_$ch.json 2> $O/wd_k_$ch.err || { echo "som_prof failed"; tail -5 $O/wd_k_$ch.err; continue; }
  python - "$ch" "$O/wd_k_$ch.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[2]))
print("chunk", sys.argv[1], "k_window_diff_ms", d["extract_tumor"]["k_window_diff_ms"], "items", d["extract_tumor"]["window_items"], "extract_tumor_ms", d["extract_tumor"]["device_ms"])
PY
done
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_k.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest_k.log
