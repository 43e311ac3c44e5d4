"""Runs only the BGZF inflation leg of bench.py (for ncu captures and A/B runs):  python tools/bgzf_prof.py [MB]"""
import importlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.entry.load_package()
import torch  # noqa: E402

host = importlib.import_module("longphase_s_b200.host")
ffi = importlib.import_module("longphase_s_b200._ffi")
args = types.SimpleNamespace(warmup=1, steps=int(os.environ.get("STEPS", "3")))
ctx = host.Context(0)
print(json.dumps(bench.bench_bgzf(ctx, host, ffi, torch, len(os.sched_getaffinity(0)), args, mb=int(sys.argv[1]) if len(sys.argv) > 1 else 256)))
