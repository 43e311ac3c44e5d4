#!/bin/bash
# First GPU call of a round (gpurun --timeout 900 -- 'bash tools/round_first_call.sh r02a'): everything the next decisions need, cheapest first,
# each step with its own timeout so that one slow step cannot eat the call.  Writes into gpurun_out/.
#   1. -m gpu tests of the host binaries (file-level parity with the reference binary)      ~1 min
#   2. file-level end to end with the opt-in paths (device inflate, deflate level)          ~2-4 min at 2 x 32 Mb
#   3. one default bench line                                                               ~2 min
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
(time timeout 150 python -m pytest tests/test_host_cli.py tests/test_host_somatic_cli.py -m gpu -x -q) > $O/host_gpu_$TAG.log 2>&1
echo "host gpu tests rc=$?" | tee -a $O/host_gpu_$TAG.log
(time timeout 420 python tools/cli_e2e.py --contigs 2 --mb ${CLI_MB:-32} --variants) > $O/cli_e2e_$TAG.json 2> $O/cli_e2e_$TAG.err
echo "cli_e2e rc=$?" | tee -a $O/cli_e2e_$TAG.err
cat $O/cli_e2e_$TAG.json
(time timeout 300 python bench.py) > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?" | tee -a $O/bench_$TAG.err
tail -c 600 $O/bench_$TAG.json
