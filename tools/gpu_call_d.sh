#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_edge_cases.py tests/test_gpu_large.py -m gpu -x -q) > $O/pytest_d.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest_d.log
(LPS_DEBUG_SLOW=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths ${BENCH_ARGS:-}) > $O/bench_d.json 2> $O/bench_d.err
echo "bench rc=$?"; grep -c "slow path" $O/bench_d.err; grep "slow path" $O/bench_d.err | sort | uniq -c | head -5; tail -2 $O/bench_d.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_d.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stage_ms"])
except Exception as e:
    print("no bench line:", e)
PY
