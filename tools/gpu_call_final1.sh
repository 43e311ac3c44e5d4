#!/bin/bash
# round-end evidence on one GPU: profile_round, then the default bench, the weak bench and the reference arm (each plain)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
TAG=${1:-r02}
bash tools/profile_round.sh $TAG > $O/profile_round_$TAG.log 2>&1; echo "profile_round rc=$?"
(time timeout 900 python bench.py) > $O/bench_${TAG}_1gpu.json 2> $O/bench_${TAG}_1gpu.err; echo "bench rc=$?"
(timeout 600 python bench.py --workload weak --no-cpu-baseline) > $O/bench_${TAG}_weak_1gpu.json 2> $O/bench_${TAG}_weak_1gpu.err; echo "weak rc=$?"
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("bench_${TAG}_1gpu", "bench_${TAG}_weak_1gpu", "bench_${TAG}_reference"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "scaling")}, "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
        if "roofline" in d: print("  roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["traffic"])
    except Exception as e:
        print(f, "no bench line:", e)
PY
