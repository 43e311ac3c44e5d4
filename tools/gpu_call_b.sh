#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(timeout 600 python -m pytest tests/test_edge_cases.py tests/test_gpu_parity.py tests/test_tag.py tests/test_tag_edge_cases.py tests/test_somatic.py -m gpu -x -q) > $O/pytest_b.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest_b.log
bash tools/ab_k1.sh "$@" 2>&1 | tee $O/ab_k1.log
