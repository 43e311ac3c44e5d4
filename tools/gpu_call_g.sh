#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_g.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest_g.log
(timeout 400 python bench.py --workload weak --no-cpu-baseline --no-other-paths) > $O/bench_g_weak.json 2> $O/bench_g_weak.err
echo "bench weak rc=$?"; tail -2 $O/bench_g_weak.err
python - <<PY
import json
for f in ("bench_g_weak",):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "scaling")}, "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
        print("  roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], [(r["kernel"], round(r["frac"], 3), round(r["kernel_ms"], 3)) for r in d["rooflines"]])
        print("  stage", d["stage_ms"]); print("  clocks", d["clocks"])
    except Exception as e:
        print(f, "no bench line:", e)
PY
bash tools/gpu_call_launches.sh r02c 2>&1 | head -24
