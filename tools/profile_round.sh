#!/bin/bash
# The captures behind profiles/: run on the GPU box (gpurun -- 'bash tools/profile_round.sh r01c'); writes into gpurun_out/.
# Every profiled command first runs once WITHOUT ncu and must exit 0.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
ONE="--steps 2 --warmup 1 --no-cpu-baseline --no-other-paths --contigs-per-gpu 1"
python bench.py $ONE > $O/plain_$TAG.json 2> $O/plain_$TAG.err || { echo "plain bench failed"; tail -5 $O/plain_$TAG.err; exit 1; }
# 1. launch list of one resident step
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv python bench.py $ONE > $O/ncu_launches_$TAG.log 2>&1
# 2. full capture of the two dominant kernels of the phase step
ncu --set full --clock-control none --import-source on -k regex:"k_call_alleles|k_fold_edges" -c 2 -f -o $O/prof_phase_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-paths --contigs-per-gpu 1 > $O/ncu_phase_$TAG.log 2>&1
ncu -i $O/prof_phase_$TAG.ncu-rep --page raw --csv > $O/phase_raw_$TAG.csv 2>/dev/null
# 3. BGZF inflation
python tools/bgzf_prof.py 256 > $O/bgzf_plain_$TAG.json 2> $O/bgzf_plain_$TAG.err || { echo "plain bgzf failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_bgzf -c 1 -f -o $O/prof_bgzf_$TAG python tools/bgzf_prof.py 256 > $O/ncu_bgzf_$TAG.log 2>&1
ncu -i $O/prof_bgzf_$TAG.ncu-rep --page raw --csv > $O/bgzf_raw_$TAG.csv 2>/dev/null
ls -la $O | grep $TAG
