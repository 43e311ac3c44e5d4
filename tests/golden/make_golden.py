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


def fingerprint(c):
    return np.array([c.n_reads, c.n_var, int(c.cigar.sum() % (1 << 31)), int(c.qual.astype(np.uint64).sum() % (1 << 31)),
                     int(c.var_pos.astype(np.int64).sum() % (1 << 31))], np.int64)


SOM_FIELDS = ["tum_var", "category", "read_hp", "ps", "pq", "h1", "h2", "h3", "n_ps", "end_pos", "read_len", "pos_base", "read_hp_count",
              "somatic_read_hp_count", "case_count", "window_hist", "hp_before_count", "hp_after_count", "h3_before_count", "h3_after_count",
              "cover_start", "cover_end", "ratios_f", "ratios_d", "case_read_count", "call_off", "calls"]


def tag_family():
    """Germline haplotag, the three somatic passes and the purity estimate, from the reference's own objects."""
    from tests import somatic_cases, tag_cases, test_purity
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kind in (("snp_indel", "phase_result"), ("many_supplementary", "blocks50")):
        c = tag_cases.get(name, kind)
        out = {"input_fingerprint": fingerprint(c)}
        for pname, tp in tag_cases.param_sets().items():
            ref = po.ReferenceTag(c, tp)
            for k in ("category", "hp", "ps", "pq", "h1", "h2", "n_ps", "var_off", "var_pos", "var_hp"):
                out[f"{pname}_{k}"] = getattr(ref, k)
            out[f"{pname}_stats"] = np.array([ref.stats[k] for k in sorted(ref.stats)], np.int64)
        path = os.path.join(here, f"tag_{name}_{kind}.npz")
        np.savez_compressed(path, **out)
        print("tag", name, kind, os.path.getsize(path) // 1024, "KiB")
    for name in ("snv_indel", "dense_somatic"):
        un, ut = somatic_cases.get(name)
        out = {"fingerprint_normal": fingerprint(un), "fingerprint_tumor": fingerprint(ut)}
        for pname in ("purity_q20", "tag_q1"):
            tp = somatic_cases.param_sets()[pname]
            for mode in ("extract_normal", "extract_tumor", "somatic_tag"):
                ref = po.ReferenceSomatic(un if mode == "extract_normal" else ut, tp, mode)
                assert ref.rc == 0
                for k in SOM_FIELDS:
                    out[f"{pname}_{mode}_{k}"] = getattr(ref, k)
                if mode == "somatic_tag":
                    out[f"{pname}_{mode}_stats"] = np.array([ref.stats[k] for k in sorted(ref.stats)], np.int64)
        path = os.path.join(here, f"somatic_{name}.npz")
        np.savez_compressed(path, **out)
        print("somatic", name, os.path.getsize(path) // 1024, "KiB")
    out = {}
    tp = somatic_cases.param_sets()["purity_q20"]
    for name in test_purity.PURITY_CASES:
        un, ut = test_purity.pair(name)
        ref = po.ReferencePurity(un, ut, tp)
        out[f"{name}_fingerprint"] = np.concatenate([fingerprint(un), fingerprint(ut)])
        out[f"{name}_purity"] = np.array([ref.purity, ref.result["median"], ref.result["q1"], ref.result["q3"], ref.result["iqr"],
                                          ref.result["lower_whisker"], ref.result["upper_whisker"]], np.float64)
        out[f"{name}_counts"] = np.array([ref.result["threshold"], ref.result["n_after_lcvf"], ref.result["n_used"]], np.int64)
    path = os.path.join(here, "purity.npz")
    np.savez_compressed(path, **out)
    print("purity", os.path.getsize(path) // 1024, "KiB")


CALL_FIELDS = ["touched", "mean_alt", "z_score", "interval_snp_count", "min_distance", "in_dense", "dense_alt_same", "filtered_by",
               "is_filter_out", "high_con", "derive_hp", "is_somatic", "flag_derive_hp", "read_hp", "read_h3"]


def somatic_call():
    """The calling stage between the extract passes and the tagging pass (SomaticVarCaller::variantCalling + getSomaticFlag), from the
    reference's own private stages run on its own extract passes."""
    from tests import somatic_cases, test_purity, test_somatic_call
    here = os.path.dirname(os.path.abspath(__file__))
    tp = somatic_cases.param_sets()["purity_q20"]
    out = {}
    for name, purity, enable_filter in test_somatic_call.CALL_CASES:
        un, ut = test_purity.pair(name)
        ref = po.ReferenceSomaticCall(un, ut, tp, purity, enable_filter)
        assert ref.rc == 0
        key = f"{name}_{purity}_{int(enable_filter)}"
        out[f"{key}_fingerprint"] = np.concatenate([fingerprint(un), fingerprint(ut)])
        out[f"{key}_tier"] = np.array([ref.tier], np.int32)
        for k in CALL_FIELDS:
            out[f"{key}_{k}"] = getattr(ref, k)
    path = os.path.join(here, "somatic_call.npz")
    np.savez_compressed(path, **out)
    print("somatic_call", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "call":
        somatic_call()
        sys.exit(0)
    if len(sys.argv) < 2 or sys.argv[1] != "tag":
        main()
    tag_family()
    somatic_call()
