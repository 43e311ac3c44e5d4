#!/bin/bash
# GPU tests, then the default bench with the 8-bit CIGAR wire format and, for comparison, the 16-bit one (end-to-end leg only differs)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/pytest_h.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest_h.log
(timeout 600 python bench.py --no-cpu-baseline --no-other-paths) > $O/bench_h8.json 2> $O/bench_h8.err
echo "bench c8 rc=$?"; tail -2 $O/bench_h8.err
(timeout 600 python bench.py --no-cpu-baseline --no-other-paths --cigar16) > $O/bench_h16.json 2> $O/bench_h16.err
echo "bench c16 rc=$?"; tail -2 $O/bench_h16.err
python - <<PY
import json
for f in ("bench_h8", "bench_h16"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "scaling")}, "e2e", d.get("e2e"))
    except Exception as e:
        print(f, "no bench line:", e)
PY
