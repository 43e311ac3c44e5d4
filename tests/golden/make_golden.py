"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_tap.so, built from
/root/reference by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Each fixture stores the synthetic-generator parameters (the inputs are regenerated deterministically from
them — libsynth is a pure function of its parameters) plus every stage the reference tap exposes:
  stage A/B/C calls, clip map, CNV intervals, sparse float edge cells, node types, sweep result, final
  result, per-alignment read haplotypes and exportResult.
tests/test_oracle_golden.py checks the oracle against them without needing the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests import cases  # noqa: E402

GOLDEN = ["snp_only", "snp_indel", "dense_indel_noseq", "pacbio_like", "many_supplementary", "short_reads_sparse"]


def flatten(prefix, d, out):
    for k, v in d.items():
        out[f"{prefix}_{k}"] = v


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name in GOLDEN:
        contig, params = cases.get(name)
        ref = po.ReferencePhase(contig, params)
        assert ref.rc == 0 and ref.complete, name
        out = {}
        flatten("a", ref.stage_a, out)
        flatten("b", ref.stage_b, out)
        flatten("c", ref.stage_c, out)
        flatten("sweep", ref.nodes_sweep, out)
        flatten("final", ref.nodes_final, out)
        for k in ("clip_pos", "clip_front", "clip_back", "cnv", "cell_a", "cell_b", "cell_which", "cell_val", "read_hp",
                  "res_pos", "res_block", "res_hap_ref", "res_hap_alt"):
            out[k] = getattr(ref, k)
        out["n_contrib"] = np.array([ref.n_contrib], np.uint64)
        out["n_empty_after_filter"] = np.array([ref.n_empty_after_filter], np.int32)
        # fingerprint of the regenerated inputs, so a generator change cannot silently invalidate the fixture
        out["input_fingerprint"] = np.array([contig.n_reads, contig.n_var, int(contig.cigar.sum() % (1 << 31)),
                                             int(contig.qual.astype(np.uint64).sum() % (1 << 31)),
                                             int(contig.var_pos.astype(np.int64).sum() % (1 << 31))], np.int64)
        path = os.path.join(here, f"phase_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB", "cells", len(ref.cell_val), "calls", len(ref.stage_a["pos"]))


if __name__ == "__main__":
    main()
