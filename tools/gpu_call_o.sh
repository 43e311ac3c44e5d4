#!/bin/bash
# GPU tests (both sweep modes), then the launch list of one step (per-kernel times) on the current build
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_o.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest_o.log
(LPS_SWEEP_SEQUENTIAL=1 timeout 900 python -m pytest tests -m gpu -x -q -k "parity or edge or large") > $O/pytest_o_seq.log 2>&1
echo "pytest (sequential sweep) rc=$?"; tail -2 $O/pytest_o_seq.log
bash tools/gpu_call_launches.sh ${1:-r02e} 2>&1 | head -40
