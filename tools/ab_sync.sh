#!/bin/bash
# Host-thread waiting modes with few cores per GPU (what each rank of an 8-GPU run on a 32-core box gets), emulated on one GPU by
# pinning the process to 4 cores:  tools/ab_sync.sh
cd "$(dirname "$0")/.."
for cfg in "--contigs-per-gpu 4 --sync spin" "--contigs-per-gpu 8 --sync yield" "--contigs-per-gpu 8 --sync block" "--contigs-per-gpu 12 --sync block"; do
  taskset -c 0-3 python bench.py --no-cpu-baseline --no-other-paths --steps 4 $cfg > /tmp/o.json 2>/tmp/o.err || { tail -3 /tmp/o.err; continue; }
  python -c "
import json; j=json.load(open('/tmp/o.json')); print('4 cores, $cfg:', 'value %.1fM step %.2f ms (%.2f ms per contig) e2e %.2fM' % (j['value']/1e6, j['ms_per_step'], j['ms_per_step']/j['config']['contigs_per_gpu'], j['e2e']['value']/1e6))"
done
