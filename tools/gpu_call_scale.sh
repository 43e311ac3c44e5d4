#!/bin/bash
# default bench at N GPUs of one box: tools/gpu_call_scale.sh 8 [extra bench args]
set -u
cd "$(dirname "$0")/.."
N=${1:-8}; shift
O=gpurun_out; mkdir -p $O
{ nproc; free -g | head -2; lscpu | grep -E "Socket|NUMA"; nvidia-smi topo -m | head -14; } > $O/box_scale_$N.txt 2>&1
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@") > $O/scale_$N.json 2> $O/scale_$N.err
echo "bench N=$N rc=$?"; tail -3 $O/scale_$N.err
python - <<PY
import json
try:
    d = json.loads(open("$O/scale_$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "scaling", "n_gpus")}, "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
    print("  stage", d["stage_ms"]); print("  cfg", {k: d["config"][k] for k in ("contigs_rank0", "host_threads_per_rank", "reads_total", "largest_rank_share_of_reads", "host_cores", "host_sync", "synth_seconds", "wall_ms_per_step_rank0")})
except Exception as e:
    print("no bench line:", e)
PY
cat $O/box_scale_$N.txt | head -8
