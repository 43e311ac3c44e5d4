#!/bin/bash
# Round-end evidence on one GPU inside a fixed time budget (seconds, $1): the GPU suite, the default bench line, then the captures
# behind profiles/ in order of importance; a stage that no longer fits is skipped and says so.
set -u
cd "$(dirname "$0")/.."
LIMIT=${1:-840}; TAG=${2:-r02}
O=gpurun_out; mkdir -p $O
left() { echo $(( LIMIT - SECONDS )); }
stage() {   # stage <seconds it needs> <name> <command...>
  local need=$1 name=$2; shift 2
  if [ $(left) -lt $need ]; then echo "[$name] skipped: $(left) s left, needs $need"; return 1; fi
  local t0=$SECONDS
  timeout $(( $(left) - 5 )) "$@"; local rc=$?
  echo "[$name] rc=$rc in $(( SECONDS - t0 )) s"
  return $rc
}
stage 150 pytest bash -c "python -m pytest tests -m gpu -q > $O/pytest_$TAG.log 2>&1; rc=\$?; tail -3 $O/pytest_$TAG.log; exit \$rc"
stage 240 bench bash -c "python bench.py > $O/bench_${TAG}_1gpu.json 2> $O/bench_${TAG}_1gpu.err"
PROFILE_STAGES="1 2" stage 150 profile_phase bash tools/profile_round.sh $TAG
PROFILE_STAGES="3c" stage 200 profile_traffic bash tools/profile_round.sh $TAG
stage 70 weak bash -c "python bench.py --workload weak --no-cpu-baseline > $O/bench_${TAG}_weak_1gpu.json 2> $O/bench_${TAG}_weak_1gpu.err"
PROFILE_STAGES="3" stage 130 profile_somatic bash tools/profile_round.sh $TAG
stage 100 reference bash -c "python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err"
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
for f in ("bench_%s_1gpu" % tag, "bench_%s_weak_1gpu" % tag, "bench_%s_reference" % tag):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "scaling")}, "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
        if "roofline" in d: print("  roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["traffic"])
    except Exception as e:
        print(f, "no bench line:", e)
PY
echo "total $SECONDS s"
