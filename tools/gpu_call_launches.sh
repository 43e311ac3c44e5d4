#!/bin/bash
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02b}
O=gpurun_out; mkdir -p $O
ONE="--workload weak --contigs-per-gpu 1 --steps 2 --warmup 1 --no-cpu-baseline --no-other-paths --no-e2e"
timeout 300 python bench.py $ONE > $O/plain_$TAG.json 2> $O/plain_$TAG.err || { echo "plain bench failed"; tail -5 $O/plain_$TAG.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_$TAG.csv python bench.py $ONE > $O/ncu_launches_$TAG.log 2>&1
echo "ncu rc=$?"
python tools/summarize_ncu.py launches $O/launches_$TAG.csv k_prep_reads 2 | head -60
