#!/bin/bash
# ncu --set full of selected kernels of one resident phase step (one contig in flight); the plain run goes first
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02}
PAT=${2:-k_call_alleles}
COUNT=${3:-2}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-paths --contigs-per-gpu 1 ${BENCH_ARGS:-}"
timeout 300 $CMD > $O/plain_$TAG.json 2> $O/plain_$TAG.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -c $COUNT -o $O/prof_$TAG -f $CMD > $O/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
tail -5 $O/ncu_$TAG.log
ls -la $O/prof_$TAG.ncu-rep
