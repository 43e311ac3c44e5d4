#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 600 python -m pytest tests/test_somatic.py tests/test_gpu_large.py tests/test_somatic_call.py tests/test_host_somatic_cli.py -m gpu -x -q) > $O/pytest_f.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest_f.log
timeout 300 python tools/som_prof.py 32 4 | tee $O/som_f.json
