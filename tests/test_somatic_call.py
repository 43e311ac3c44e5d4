"""The calling stage of somatic_haplotag (lps_somatic_call = SomaticVarCaller::variantCalling without extraction / logs, plus
getSomaticFlag) against the UNMODIFIED reference run on its own extract passes.  CPU: the extract products come from the oracle;
-m gpu: from the CUDA path, and the flags then drive the tagging pass."""
import importlib

import numpy as np
import pytest

from . import somatic_cases
from .test_purity import PURITY_CASES, pair

po = pytest.importorskip("oracle.pyoracle")
host = importlib.import_module("longphase_s_b200.host")
ffi = importlib.import_module("longphase_s_b200._ffi")


def as_result(o):
    return {k: getattr(o, k) for k in ("tum_var", "pos_base", "read_hp_count", "somatic_read_hp_count", "ratios_f", "case_read_count",
                                       "window_hist", "category", "read_hp", "h1", "h2", "h3", "n_ps", "end_pos", "call_off", "calls")}


def compare(got, ref):
    assert ref.rc == 0
    assert got["tier"] == ref.tier
    assert np.array_equal(got["touched"], ref.touched), "touched positions differ"
    t = ref.touched.astype(bool)
    assert got["mean_alt_per_var_read"][t].tobytes() == ref.mean_alt[t].tobytes(), "meanAltCountPerVarRead differs"
    assert got["z_score"][t].tobytes() == ref.z_score[t].tobytes(), "zScore differs"
    for a, b in (("interval_snp_count", "interval_snp_count"), ("min_distance", "min_distance"), ("in_dense_interval", "in_dense"),
                 ("dense_alt_same_count", "dense_alt_same"), ("is_filter_out", "is_filter_out"), ("is_somatic", "high_con")):
        assert np.array_equal(got[a][t], getattr(ref, b)[t]), a
    assert np.array_equal(got["filtered_by"][t], ref.filtered_by[t]), "per-filter flags differ"
    assert np.array_equal(got["derive_hp"][t], ref.derive_hp[t]), "somaticReadDeriveByHP differs"
    # what getSomaticFlag leaves in the variant map, for every tumor position
    assert np.array_equal(got["is_somatic"], ref.is_somatic)
    assert np.array_equal(got["derive_hp"][ref.is_somatic.astype(bool)], ref.flag_derive_hp[ref.is_somatic.astype(bool)])
    assert np.array_equal(got["read_hp"], ref.read_hp), "calibrated read haplotypes differ"
    assert np.array_equal(got["read_h3"], ref.read_h3), "calibrated HP3 counts differ"


CALL_CASES = [("purity_60", 0.6, True), ("purity_60", 0.95, True), ("purity_60", 0.75, True), ("purity_30", 0.3, True), ("purity_30", 0.1, True),
              ("purity_90_dense", 0.9, True), ("purity_90_dense", 0.45, False), ("no_support", 0.5, True)]


@pytest.mark.skipif(not po.tap_available(), reason="reference tap not built")
@pytest.mark.parametrize("name,purity,enable_filter", CALL_CASES)
def test_somatic_call_matches_reference(name, purity, enable_filter):
    un, ut = pair(name)
    tp = somatic_cases.param_sets()["purity_q20"]
    ref = po.ReferenceSomaticCall(un, ut, tp, purity, enable_filter)
    on, ot = po.OracleSomatic(un, tp, "extract_normal"), po.OracleSomatic(ut, tp, "extract_tumor")
    got = host.SomaticVarCaller(ut, tp, enable_filter).variantCalling(as_result(on), as_result(ot), purity)
    compare(got, ref)
    if name != "no_support" and enable_filter:
        assert 0 < got["n_somatic"] < int(got["touched"].sum()), "the case should both keep and filter positions"
        assert got["filtered_by"].sum(0).astype(bool).sum() >= 3, "several filters should fire"


def test_somatic_call_rejects_bad_arguments():
    lib = ffi.load_library()
    assert lib.lps_somatic_call(None, None) == -1


@pytest.mark.gpu
def test_gpu_somatic_pipeline_matches_reference():
    """extract normal -> extract tumor -> purity -> calling -> tagging, all through the C ABI, against the reference's own chain."""
    un, ut = pair("purity_60")
    tp = somatic_cases.param_sets()["purity_q20"]
    ctx = host.Context(0)
    try:
        rn = host.ExtractNorDataChrProcessor(ctx, un, tp).processSingleChrom(un)
        rt = host.ExtractTumDataChrProcessor(ctx, ut, tp).processSingleChrom(ut)
        purity = host.TumorPurityEstimator([rn], [rt]).estimateTumorPurity()
        got = host.SomaticVarCaller(ut, tp).variantCalling(rn, rt, purity)
        on, ot = po.OracleSomatic(un, tp, "extract_normal"), po.OracleSomatic(ut, tp, "extract_tumor")
        want = host.SomaticVarCaller(ut, tp).variantCalling(as_result(on), as_result(ot), purity)
        for k in want:
            assert np.array_equal(got[k], want[k]), k
        if po.tap_available():
            ref = po.ReferenceSomaticCall(un, ut, tp, purity, True)
            compare(got, ref)
        # the flags drive the tagging pass: same tags as the oracle's tagging pass given the same flags
        import copy
        tagged = copy.copy(ut)
        tagged.is_somatic = np.zeros(ut.n_var, np.uint8)
        tagged.derive_hp = np.zeros(ut.n_var, np.int8)
        tagged.is_somatic[got["tum_var"]] = got["is_somatic"]
        tagged.derive_hp[got["tum_var"]] = got["derive_hp"]
        res = host.SomaticHaplotagChrProcessor(ctx, tagged, tp).processSingleChrom(tagged)
        otag = po.OracleSomatic(tagged, tp, "somatic_tag")
        assert np.array_equal(res["read_hp"], otag.read_hp) and np.array_equal(res["ps"], otag.ps) and np.array_equal(res["pq"], otag.pq)
        assert int((res["read_hp"] >= 3).sum()) > 0, "some reads should carry a somatic haplotype"
    finally:
        ctx.close()


def test_somatic_call_reports_the_reference_exit():
    """A tumor position that alignments reached (depth > 0) but that no alignment recorded in tumorPosReadCorrBaseHP: the reference
    prints "[ERROR](calibrate read HP) => can't find pos" and exits (SomaticVarCaller.cpp:1397-1400); the library returns LPS_E_DATA."""
    import types
    nt, nr = 1, 1
    pb = np.zeros((nt, 15), np.int32)
    pb[0, 6] = 5                                               # LPS_PB_DEPTH: only low-MAPQ / deletion coverage
    z9 = np.zeros((nt, 9), np.int32)
    res = dict(tum_var=np.array([0], np.int32), pos_base=pb, read_hp_count=z9, somatic_read_hp_count=z9, ratios_f=np.zeros((nt, 9), np.float32),
               case_read_count=np.zeros(nt, np.int32), window_hist=np.zeros((nt, 2, 201), np.int32), category=np.zeros(nr, np.uint8),
               read_hp=np.zeros(nr, np.int8), h1=np.zeros(nr, np.int32), h2=np.zeros(nr, np.int32), h3=np.zeros(nr, np.int32),
               n_ps=np.zeros(nr, np.uint8), end_pos=np.zeros(nr, np.int32), call_off=np.zeros(nr + 1, np.uint64), calls=np.zeros(0, ffi.CALL_DTYPE))
    contig = types.SimpleNamespace(var_pos=np.array([100], np.int32), tum_ref_len=np.array([1], np.uint16), tum_alt_len=np.array([1], np.uint16))
    tp = somatic_cases.param_sets()["purity_q20"]
    with pytest.raises(host.LpsError) as e:
        host.SomaticVarCaller(contig, tp).variantCalling(res, res, 0.6)
    assert e.value.code == -6
    # with the filters disabled the position counts as somatic and the reference exits in statisticSomaticPosReadHP instead (:1511-1514)
    with pytest.raises(host.LpsError):
        host.SomaticVarCaller(contig, tp, enableFilter=False).variantCalling(res, res, 0.6)
