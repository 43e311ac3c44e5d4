#!/usr/bin/env python
"""Profiling driver of the somatic dialects (BASELINE config C4 shard): one tumor 50x / normal 25x pair over a 32 Mb contig through
lps_extract_normal, lps_extract_tumor (k_call_alleles<EXTRACT_TUMOR> + k_window_diff) and lps_somatic_tag_reads, a few times each.
Prints one JSON line with the device times; run it under ncu to capture the kernels (tools/profile_round.sh)."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
synth = importlib.import_module("longphase_s_b200.synth")
host = importlib.import_module("longphase_s_b200.host")
ffi = importlib.import_module("longphase_s_b200._ffi")
wl = importlib.import_module("longphase_s_b200.workloads")

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 32.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
un, ut = wl.c4_pair(synth, mb)
sp = ffi.LpsTagParams(mapping_quality=20, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6)
ctx = host.Context(0)
out = {"contig_mb": mb, "normal_reads": int(un.n_reads), "tumor_reads": int(ut.n_reads), "positions": int(un.n_var)}
for name, c, cls in (("extract_normal", un, host.ExtractNorDataChrProcessor), ("extract_tumor", ut, host.ExtractTumDataChrProcessor),
                     ("somatic_tag", ut, host.SomaticHaplotagChrProcessor)):
    proc = cls(ctx, c, sp)
    ms, k1, wd = [], [], []
    for _ in range(reps):
        r = proc.processSingleChrom(c)
        st = ctx.stats()
        ms.append(st["ms_tag_reads"]); k1.append(st["ms_kernel_call_alleles"]); wd.append(st["ms_kernel_window_diff"])
    out[name] = {"device_ms": float(np.mean(ms[1:] or ms)), "k_call_alleles_ms": float(np.mean(k1[1:] or k1)), "n_tum": int(r["n_tum"])}
    if name == "extract_tumor":
        out[name].update(k_window_diff_ms=float(np.mean(wd[1:] or wd)), window_items=int(r["n_window_items"]))
ctx.close()
print(json.dumps(out))
