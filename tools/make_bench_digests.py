#!/usr/bin/env python
"""Writes tests/golden/bench_digests.json: the oracle's digest (workloads.phase_digest over ps / hap_ref / read_hp / hp_counts) of
every contig bench.py times by default, keyed by the synth configuration.  bench.py compares the digest of what the GPU produced
with these and refuses to print a value when one differs (SURVEY 8d "correctness gate before
timing").  The oracle (oracle/, test infrastructure) runs HERE, never inside bench.py's timed path.

usage: python tools/make_bench_digests.py [--ranks 8] [--contigs 8] [--genome-mb 1024 ...] [--jobs 8]
"""
import argparse
import importlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(kw):
    import __graft_entry__ as entry
    entry.load_package()
    synth = importlib.import_module("longphase_s_b200.synth")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    wl = importlib.import_module("longphase_s_b200.workloads")
    from oracle import pyoracle as po
    from tests import parity
    c = synth.Contig(**kw)
    orc = po.OraclePhase(c, ffi.default_phase_params(True))
    return wl.key_of(kw), {"digest": parity.oracle_phase_digest(orc, c.n_reads), "reads": int(c.n_reads), "variants": int(c.n_var),
                           "calls": int(len(orc.calls)), "phased": int((orc.ps != 0).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--contigs", type=int, default=8)
    ap.add_argument("--genome-mb", type=float, nargs="*", default=[])
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) // 2))
    args = ap.parse_args()
    import __graft_entry__ as entry
    entry.load_package()
    wl = importlib.import_module("longphase_s_b200.workloads")
    have = wl.load_digests()
    todo = []
    for r in range(args.ranks):
        for i in range(args.contigs):
            todo.append(wl.phase_kwargs(wl.weak_seed(r, i)))
    for g in args.genome_mb:
        for _, seed, mb in wl.genome_contigs(g):
            todo.append(wl.phase_kwargs(seed, mb))
    todo = [kw for kw in todo if wl.key_of(kw) not in have]
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    with ProcessPoolExecutor(max_workers=args.jobs) as ex:
        for k, v in ex.map(one, todo):
            have[k] = v
            print(k, v, flush=True)
    os.makedirs(os.path.dirname(wl.DIGEST_FILE), exist_ok=True)
    with open(wl.DIGEST_FILE, "w") as f:
        json.dump(have, f, indent=0, sort_keys=True)
    print("wrote", wl.DIGEST_FILE, len(have), "entries")


if __name__ == "__main__":
    main()
