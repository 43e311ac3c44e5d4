#!/bin/bash
# GPU call: box facts, smoke, the -m gpu suite, one short bench line (each step under its own timeout)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
{ nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; nvidia-smi --query-gpu=index,name,memory.total --format=csv; nvidia-smi topo -m 2>/dev/null | head -20; } > $O/box_$TAG.txt 2>&1
(time timeout 180 python __graft_entry__.py smoke) > $O/smoke_$TAG.log 2>&1
echo "smoke rc=$?" | tee -a $O/smoke_$TAG.log
tail -8 $O/smoke_$TAG.log
(time timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-}) > $O/pytest_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log
tail -30 $O/pytest_$TAG.log
(time LPS_DEBUG_K1=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-paths ${BENCH_ARGS:-}) > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?" | tee -a $O/bench_$TAG.err
grep -E "^k1 " $O/bench_$TAG.err | tail -4
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"], d["stage_ms"])
except Exception as e:
    print("no bench line:", e)
PY
