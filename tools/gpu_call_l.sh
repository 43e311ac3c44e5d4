#!/bin/bash
# somatic GPU tests + C4 shard timing + ncu of k_window_diff on the current build
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TAG=${1:-r02c}
(time timeout 900 python -m pytest tests -m gpu -x -q -k "somatic or c4 or window or purity") > $O/pytest_l.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest_l.log
bash tools/gpu_call_j.sh $TAG 2>&1 | grep -v "^$" | head -30
