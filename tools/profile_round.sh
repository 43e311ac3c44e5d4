#!/bin/bash
# The captures behind profiles/: run on the GPU box (gpurun -- 'bash tools/profile_round.sh r02'); writes into gpurun_out/.
# Every profiled command first runs once WITHOUT ncu and must exit 0.  One contig in flight, so that launches do not overlap.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
# PROFILE_STAGES picks the captures (default: all): 1 launch list, 2 phase kernels, 3 somatic dialects + window diff, 3c DRAM traffic of the genome bench
ST=" ${PROFILE_STAGES:-1 2 3 3c} "
has() { case "$ST" in *" $1 "*) return 0;; *) return 1;; esac; }
ONE="--workload weak --contigs-per-gpu 1 --steps 2 --warmup 1 --no-cpu-baseline --no-other-paths --no-e2e"
# 1. launch list of one resident step
if has 1 || has 2; then
timeout 300 python bench.py $ONE > $O/plain_$TAG.json 2> $O/plain_$TAG.err || { echo "plain bench failed"; tail -5 $O/plain_$TAG.err; exit 1; }
fi
if has 1; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_$TAG.csv python bench.py $ONE > $O/ncu_launches_$TAG.log 2>&1
fi
# 2. full capture of the kernels of the phase step that matter
if has 2; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_call_alleles|k_fold_edges|k_sweep_segments|k_read_vote|k_prep_reads" -s 10 -c 8 -f -o $O/prof_phase_$TAG \
    python bench.py $ONE > $O/ncu_phase_$TAG.log 2>&1
ncu -i $O/prof_phase_$TAG.ncu-rep --page raw --csv > $O/phase_raw_$TAG.csv 2>/dev/null
fi
# 3. the somatic dialects (C4 shard): k_call_alleles<2,3,4>, k_window_diff
if has 3; then
timeout 300 python tools/som_prof.py 32 3 > $O/som_plain_$TAG.json 2> $O/som_plain_$TAG.err || { echo "plain somatic failed"; tail -5 $O/som_plain_$TAG.err; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_call_alleles|k_window_diff" -c 4 -f -o $O/prof_som_$TAG python tools/som_prof.py 32 1 > $O/ncu_som_$TAG.log 2>&1
ncu -i $O/prof_som_$TAG.ncu-rep --page raw --csv > $O/som_raw_$TAG.csv 2>/dev/null
# 3b. k_window_diff alone (the C4 shard again)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_window_diff" -c 1 -f -o $O/prof_wd_$TAG python tools/som_prof.py 32 1 > $O/ncu_wd_$TAG.log 2>&1
ncu -i $O/prof_wd_$TAG.ncu-rep --page raw --csv > $O/wd_raw_$TAG.csv 2>/dev/null
fi
# 3c. DRAM traffic of every k_call_alleles launch of the default (genome) bench: the largest one is the launch bench.py's roofline is quoted on
GEN="--steps 1 --warmup 1 --no-cpu-baseline --no-other-paths --no-e2e"
if has 3c; then
timeout 300 python bench.py $GEN > $O/plain_genome_$TAG.json 2> $O/plain_genome_$TAG.err || { echo "plain genome bench failed"; tail -5 $O/plain_genome_$TAG.err; exit 1; }
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_call_alleles" -c 200 --csv \
    --log-file $O/k1_traffic_genome_$TAG.csv python bench.py $GEN > $O/ncu_k1_traffic_$TAG.log 2>&1
fi
# 4. BGZF inflation
if [ "${PROFILE_BGZF:-0}" = "1" ]; then
  python tools/bgzf_prof.py 256 > $O/bgzf_plain_$TAG.json 2> $O/bgzf_plain_$TAG.err || { echo "plain bgzf failed"; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:k_bgzf -c 1 -f -o $O/prof_bgzf_$TAG python tools/bgzf_prof.py 256 > $O/ncu_bgzf_$TAG.log 2>&1
  ncu -i $O/prof_bgzf_$TAG.ncu-rep --page raw --csv > $O/bgzf_raw_$TAG.csv 2>/dev/null
fi
ls -la $O | grep $TAG
