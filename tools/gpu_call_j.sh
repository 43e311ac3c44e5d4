#!/bin/bash
# ncu --set full of k_window_diff (the C4 shard of tools/som_prof.py), after the same command ran plain and exited 0
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TAG=${1:-r02}
timeout 300 python tools/som_prof.py 32 3 > $O/wd_plain_$TAG.json 2> $O/wd_plain_$TAG.err || { echo "plain somatic failed"; tail -5 $O/wd_plain_$TAG.err; exit 1; }
cat $O/wd_plain_$TAG.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_window_diff" -c 1 -f -o $O/prof_wd_$TAG python tools/som_prof.py 32 1 > $O/ncu_wd_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i $O/prof_wd_$TAG.ncu-rep --page raw --csv > $O/wd_raw_$TAG.csv 2>/dev/null
python tools/summarize_ncu.py raw $O/wd_raw_$TAG.csv
