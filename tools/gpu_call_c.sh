#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_c.log 2>&1
echo "pytest rc=$?"; tail -25 $O/pytest_c.log
(LPS_SWEEP_SEQUENTIAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_edge_cases.py -m gpu -x -q) > $O/pytest_c_seq.log 2>&1
echo "pytest sequential-sweep rc=$?"; tail -3 $O/pytest_c_seq.log
(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths) > $O/bench_c.json 2> $O/bench_c.err
echo "bench rc=$?"; tail -2 $O/bench_c.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_c.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["stage_ms"])
except Exception as e:
    print("no bench line:", e)
PY
bash tools/ab_k1.sh "$@" 2>&1 | tee $O/ab_k1_c.log
