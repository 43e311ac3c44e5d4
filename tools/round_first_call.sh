#!/bin/bash
# First GPU call of a round (gpurun --timeout 900 -- 'bash tools/round_first_call.sh r03a'): what round 2 could not measure for lack of
# GPU minutes, cheapest first, each step with its own timeout so that one slow step cannot eat the call.  Writes into gpurun_out/.
#   1. the whole -m gpu suite (the device BAM writer is the default of the product binaries since r02)            ~1.5 min
#   2. file level at the size of BASELINE config C1 (one 64 Mb contig, ~4 GB BAM): reference binary, this host with the
#      device writer, with htslib's writer, with four slice readers                                              ~4-6 min
#   3. one default bench line (now with the k_bgzf_deflate roofline entry)                                        ~3 min
# On a box with several GPUs afterwards: bash tools/gpu_call_scale.sh (config.numa_bind in the line says what the binding did).
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r03a}
O=gpurun_out
mkdir -p $O
(time timeout 200 python -m pytest tests -m gpu -q) > $O/gpu_suite_$TAG.log 2>&1
echo "gpu suite rc=$?" | tee -a $O/gpu_suite_$TAG.log
tail -3 $O/gpu_suite_$TAG.log
(time timeout 480 python tools/cli_e2e.py --contigs 1 --mb ${CLI_MB:-64} --deflate) > $O/cli_e2e_$TAG.json 2> $O/cli_e2e_$TAG.err
echo "cli_e2e rc=$?" | tee -a $O/cli_e2e_$TAG.err
cat $O/cli_e2e_$TAG.json
(time timeout 300 python bench.py) > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?" | tee -a $O/bench_$TAG.err
tail -c 600 $O/bench_$TAG.json
