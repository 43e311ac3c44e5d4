#!/usr/bin/env python
"""Per-source-line totals of an `ncu --page source --csv --print-source cuda,sass` export: instructions executed and stall samples,
aggregated over the SASS of each CUDA line (first launch in the file).  usage: ncu_source_lines.py <csv> [top N] [file substring]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors="replace")))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
# two "Source" columns: CUDA source text and SASS text
i_line, i_src, i_sass = 0, 1, 3
i_inst, i_samp, i_ni = col["Instructions Executed"], col["# Samples"], col["Warp Stall Sampling (Not-issued Samples)"]
inst, samp, text = defaultdict(float), defaultdict(float), {}
tot_i = tot_s = 0.0
end = next((i for i in range(hdr_i + 1, len(rows)) if rows[i] and rows[i][0] == "File Path"), len(rows))
for r in rows[hdr_i + 1:end]:
    if len(r) <= i_inst:
        continue
    try:
        ln = int(r[i_line])
    except ValueError:
        continue
    try:
        a = float(r[i_inst] or 0)
        s = float(r[i_samp] or 0)
    except ValueError:
        continue
    inst[ln] += a; samp[ln] += s; tot_i += a; tot_s += s
    if r[i_src]:
        text[ln] = r[i_src].strip()
print(f"total warp instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
print("by instructions:")
for ln, v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{ln:5d} {100 * v / tot_i:5.1f}% inst {100 * samp[ln] / max(tot_s, 1):5.1f}% samples | {text.get(ln, '')[:110]}")
print("by samples:")
for ln, v in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{ln:5d} {100 * inst[ln] / tot_i:5.1f}% inst {100 * v / max(tot_s, 1):5.1f}% samples | {text.get(ln, '')[:110]}")
