"""Turns ncu CSV exports into the markdown summaries kept under profiles/.
  python tools/summarize_ncu.py launches <launch_list.csv> [marker kernel] [index]
        per-kernel totals and shares of ONE step: the launches from the index-th launch of the marker kernel (default: the third
        k_first_var, i.e. the last timed resident step of `bench.py --steps 2 --warmup 1`) up to the next one
  python tools/summarize_ncu.py raw <raw_page.csv> [kernel regex]    key metrics of each captured launch"""
import csv
import re
import sys
from collections import OrderedDict

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct",
       "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path, marker="k_first_var", index=2):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot = OrderedDict()
    body = [r for r in rows[1:] if r[im] == "gpu__time_duration.sum"]
    marks = [i for i, r in enumerate(body) if marker in r[ik]]
    if marker and len(marks) > index:
        body = body[marks[index]:marks[index + 1] if len(marks) > index + 1 else len(body)]
    for r in body:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[iu], 1.0)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {t:.1f} | {100 * t / total:.1f}% |")
    print(f"\nTotal GPU time under ncu: {total / 1e3:.3f} ms over {sum(n for n, _ in tot.values())} launches.")


def raw(path, pattern=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ik]):
            continue
        print(f"\n### `{re.sub(r'[(].*', '', r[ik])[:100]}`\n\n| metric | value |\n|---|---|")
        vals = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        for k in KEY:
            if k in vals:
                print(f"| {k} | {vals[k]} {u[k]} |")
        stalls = sorted(((float(v.replace(',', '')), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''))
                         for h, v in vals.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v),
                        reverse=True)[:6]
        print("| top stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], *(sys.argv[3:4] or ["k_first_var"]), *(int(a) for a in sys.argv[4:5]))
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
