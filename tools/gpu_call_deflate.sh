#!/bin/bash
# the device deflate on the GPU box: kernel == host-compiled encoder (pytest), throughput, and the tagged-BAM writer at file level
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 70 python -m pytest tests/test_bgzf_deflate.py -m gpu -q > $O/deflate_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/deflate_pytest.log
timeout 50 python tools/deflate_prof.py 256 > $O/deflate_prof.json 2> $O/deflate_prof.err; echo "prof rc=$?"; cat $O/deflate_prof.json; tail -2 $O/deflate_prof.err
timeout 120 python tools/cli_e2e.py --contigs 2 --mb 8 --deflate > $O/cli_e2e_deflate.json 2> $O/cli_e2e_deflate.err; echo "cli rc=$?"; cat $O/cli_e2e_deflate.json; tail -3 $O/cli_e2e_deflate.err
echo "total $SECONDS s"
