"""The oracle against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py).
Runs anywhere (no /root/reference, no GPU)."""
import glob
import os

import numpy as np
import pytest

from . import cases, oracle_checks

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[len("phase_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "phase_*.npz")))


def test_fixtures_exist():
    assert len(NAMES) >= 5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    from oracle import pyoracle as po
    g = np.load(os.path.join(GOLDEN_DIR, f"phase_{name}.npz"))
    contig, params = cases.get(name)
    fp = np.array([contig.n_reads, contig.n_var, int(contig.cigar.sum() % (1 << 31)),
                   int(contig.qual.astype(np.uint64).sum() % (1 << 31)), int(contig.var_pos.astype(np.int64).sum() % (1 << 31))], np.int64)
    assert np.array_equal(fp, g["input_fingerprint"]), "synthetic generator no longer reproduces the fixture's inputs"
    orc = oracle_checks.check_oracle_against(g, contig, params, po)
    assert len(orc.calls) > 0 and (orc.ps != 0).sum() > 0


# ---- tag family: germline haplotag, the three somatic passes, purity (fixtures written by make_golden.py tag) ----
def _fp(c):
    return np.array([c.n_reads, c.n_var, int(c.cigar.sum() % (1 << 31)), int(c.qual.astype(np.uint64).sum() % (1 << 31)),
                     int(c.var_pos.astype(np.int64).sum() % (1 << 31))], np.int64)


@pytest.mark.parametrize("name,kind", [("snp_indel", "phase_result"), ("many_supplementary", "blocks50")])
def test_tag_oracle_matches_reference_golden(name, kind):
    from oracle import pyoracle as po
    from . import tag_cases
    g = np.load(os.path.join(GOLDEN_DIR, f"tag_{name}_{kind}.npz"))
    c = tag_cases.get(name, kind)
    assert np.array_equal(_fp(c), g["input_fingerprint"])
    ok = c.l_qseq > 0                                               # `*` SEQ: the reference reads past the record (undefined)
    for pname, tp in tag_cases.param_sets().items():
        orc = po.OracleTag(c, tp)
        assert np.array_equal(orc.category, g[f"{pname}_category"])
        for k in ("hp", "ps", "pq", "h1", "h2"):
            assert np.array_equal(getattr(orc, k)[ok], g[f"{pname}_{k}"][ok]), (pname, k)


@pytest.mark.parametrize("name", ["snv_indel", "dense_somatic"])
def test_somatic_oracle_matches_reference_golden(name):
    from oracle import pyoracle as po
    from . import somatic_cases
    from .test_somatic import PER_SLOT
    g = np.load(os.path.join(GOLDEN_DIR, f"somatic_{name}.npz"))
    un, ut = somatic_cases.get(name)
    assert np.array_equal(_fp(un), g["fingerprint_normal"]) and np.array_equal(_fp(ut), g["fingerprint_tumor"])
    for pname in ("purity_q20", "tag_q1"):
        tp = somatic_cases.param_sets()[pname]
        for mode in ("extract_normal", "extract_tumor", "somatic_tag"):
            orc = po.OracleSomatic(un if mode == "extract_normal" else ut, tp, mode)
            key = lambda k: g[f"{pname}_{mode}_{k}"]  # noqa: E731
            assert np.array_equal(orc.category, key("category")) and np.array_equal(orc.tum_var, key("tum_var"))
            for k in PER_SLOT:
                assert np.array_equal(getattr(orc, k), key(k)), (pname, mode, k)
            proc = orc.category == 0
            if mode == "extract_tumor":
                proc = proc & (key("read_hp") != -1)
            for k in (("read_hp", "ps", "pq") if mode == "somatic_tag" else ("read_hp", "h1", "h2", "n_ps", "end_pos", "read_len")):
                assert np.array_equal(getattr(orc, k)[proc], key(k)[proc]), (pname, mode, k)


def test_purity_matches_reference_golden():
    import importlib
    from oracle import pyoracle as po
    from . import somatic_cases, test_purity
    host = importlib.import_module("longphase_s_b200.host")
    g = np.load(os.path.join(GOLDEN_DIR, "purity.npz"))
    tp = somatic_cases.param_sets()["purity_q20"]
    for name in test_purity.PURITY_CASES:
        un, ut = test_purity.pair(name)
        assert np.array_equal(np.concatenate([_fp(un), _fp(ut)]), g[f"{name}_fingerprint"])
        on, ot = po.OracleSomatic(un, tp, "extract_normal"), po.OracleSomatic(ut, tp, "extract_tumor")
        est = host.TumorPurityEstimator([test_purity.as_result(on)], [test_purity.as_result(ot)])
        purity = est.estimateTumorPurity()
        ref, cnt = g[f"{name}_purity"], g[f"{name}_counts"]
        assert purity == ref[0], (name, purity, ref[0])
        if cnt[2] >= 0:
            r = est.result
            assert [r["median"], r["q1"], r["q3"], r["iqr"], r["lower_whisker"], r["upper_whisker"]] == list(ref[1:])
            assert [r["read_count_threshold"], r["n_after_lcvf"], r["n_used"]] == list(cnt)


def test_somatic_call_matches_reference_golden():
    """lps_somatic_call on the oracle's extract passes against what the reference's own calling stage produced (fixture)."""
    import importlib
    import types
    from oracle import pyoracle as po
    from . import somatic_cases, test_purity, test_somatic_call
    host = importlib.import_module("longphase_s_b200.host")
    g = np.load(os.path.join(GOLDEN_DIR, "somatic_call.npz"))
    tp = somatic_cases.param_sets()["purity_q20"]
    passes = {}
    for name, purity, enable_filter in test_somatic_call.CALL_CASES:
        un, ut = test_purity.pair(name)
        key = f"{name}_{purity}_{int(enable_filter)}"
        assert np.array_equal(np.concatenate([_fp(un), _fp(ut)]), g[f"{key}_fingerprint"])
        if name not in passes:
            passes[name] = (po.OracleSomatic(un, tp, "extract_normal"), po.OracleSomatic(ut, tp, "extract_tumor"))
        on, ot = passes[name]
        got = host.SomaticVarCaller(ut, tp, enable_filter).variantCalling(test_somatic_call.as_result(on), test_somatic_call.as_result(ot), purity)
        ref = types.SimpleNamespace(rc=0, tier=int(g[f"{key}_tier"][0]), **{k: g[f"{key}_{k}"] for k in
                                    ("touched", "mean_alt", "z_score", "interval_snp_count", "min_distance", "in_dense", "dense_alt_same",
                                     "filtered_by", "is_filter_out", "high_con", "derive_hp", "is_somatic", "flag_derive_hp", "read_hp", "read_h3")})
        test_somatic_call.compare(got, ref)
