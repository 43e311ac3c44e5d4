#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_e.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest_e.log
(time timeout 400 python bench.py --workload weak --no-cpu-baseline --no-other-paths) > $O/bench_e_weak.json 2> $O/bench_e_weak.err
echo "bench weak rc=$?"; tail -2 $O/bench_e_weak.err
(time timeout 900 python bench.py) > $O/bench_e.json 2> $O/bench_e.err
echo "bench default rc=$?"; tail -4 $O/bench_e.err
python - <<PY
import json
for f in ("bench_e_weak", "bench_e"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "scaling")}, "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
        print("  roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], [(r["kernel"], round(r["frac"], 3), round(r["kernel_ms"], 3)) for r in d["rooflines"]])
        print("  stage", d["stage_ms"]); print("  cpu", d.get("cpu_baseline")); print("  synth_s", d["config"]["synth_seconds"], "clocks", d["clocks"])
        if "other_paths" in d: print("  other", {k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a in ("reads_per_s", "device_ms", "k_window_diff_ms", "k_window_diff_gbs", "error", "out_gb_per_s")}) for k, v in d["other_paths"].items()})
    except Exception as e:
        print(f, "no bench line:", e)
PY
free -g | head -2
