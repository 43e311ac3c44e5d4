#!/bin/bash
# round 2, last GPU call: the GPU suite with the interleaved SEQ + QUAL rows in every phase parity case, the default bench with both
# end-to-end wire formats, and the resident leg with the rows in HBM against the two arrays (k_call_alleles alone, 2 x 64 Mb)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
(time timeout 540 python -m pytest tests -m gpu -x -q) > $O/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/q_pytest.log
(time timeout 480 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --e2e-ab) > $O/q_bench.json 2> $O/q_bench.err; echo "bench rc=$?"; tail -3 $O/q_bench.err
for v in two_arrays rows; do
  EX=""; [ $v = rows ] && EX="--resident-sq"
  timeout 240 python bench.py --workload weak --contigs-per-gpu 2 --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --no-e2e $EX > $O/q_weak_$v.json 2> $O/q_weak_$v.err; echo "weak $v rc=$?"
done
python - <<'PY'
import json
def last(f):
    try: return json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e: return None
d = last("q_bench")
if d:
    print("default: ms", d["ms_per_step"], "value", d["value"], "k1 frac", d["roofline"]["frac"], "k1 ms", d["roofline"]["kernel_ms"])
    print("e2e:", {k: d["e2e"].get(k) for k in ("ms_per_step", "value", "h2d_bytes_per_step", "seq_qual_wire_format", "other_wire_format")})
for v in ("two_arrays", "rows"):
    d = last("q_weak_" + v)
    if d: print("weak", v, "ms", d["ms_per_step"], "k1 alone", d["stage_ms"]["k_call_alleles_alone"], "fold alone", d["stage_ms"]["k_fold_edges_alone"])
PY
