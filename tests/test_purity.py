"""Tumor purity (TumorPurityEstimator): lps_estimate_purity fed by the extract passes vs the unmodified reference's estimator run on
its own passes.  CPU: the extract products come from the oracle; -m gpu: from the CUDA path.  north_star asks for 1e-6 relative;
the arithmetic is restated in the reference's own types, so the values are compared for equality."""
import importlib

import numpy as np
import pytest

from . import somatic_cases

po = pytest.importorskip("oracle.pyoracle")
host = importlib.import_module("longphase_s_b200.host")
ffi = importlib.import_module("longphase_s_b200._ffi")

PURITY_CASES = {
    # deeper / longer pairs than the parity cases, so that enough positions pass the low-confidence filters
    "purity_60": (dict(seed=51, contig_len=3_000_000, indel_frac=0.1, somatic_rate=1 / 3000.0), 25, 50, 0.6, 8),
    "purity_30": (dict(seed=52, contig_len=3_000_000, indel_frac=0.1, somatic_rate=1 / 3000.0), 25, 50, 0.3, 9),
    "purity_90_dense": (dict(seed=53, contig_len=1_500_000, indel_frac=0.15, variant_rate=1 / 500.0, somatic_rate=1 / 1500.0), 30, 60, 0.9, 10, 400),
    # phase sets shorter than the reads: nearly every normal read crosses two blocks, no position passes the filters, both sides report 0.0
    "no_support": (dict(seed=54, contig_len=800_000, indel_frac=0.1, variant_rate=1 / 500.0, somatic_rate=1 / 1500.0), 30, 60, 0.5, 11, 20),
}
_cache = {}


def pair(name):
    if name not in _cache:
        kw, dn, dt, purity, useed = PURITY_CASES[name][:5]
        block = PURITY_CASES[name][5] if len(PURITY_CASES[name]) > 5 else 50
        synth = somatic_cases.synth
        cn = synth.Contig(**kw, depth=dn, purity=0.0, read_seed=3000 + kw["seed"])
        ct = synth.Contig(**kw, depth=dt, purity=purity, read_seed=4000 + kw["seed"])
        un = cn.somatic_union(seed=useed, block=block)
        _cache[name] = (un, un.with_reads_of(ct))
    return _cache[name]


def as_result(o):
    return dict(ratios_d=o.ratios_d, read_hp_count=o.read_hp_count)


def compare(est, purity, ref):
    r = est.result
    if ref.result["n_used"] < 0:                              # the reference's estimator threw: it reports purity 0.0
        assert ref.purity == 0.0 and purity == 0.0 and r["ok"] == 0
        return
    assert r["read_count_threshold"] == ref.result["threshold"]
    assert r["n_after_lcvf"] == ref.result["n_after_lcvf"] and r["n_used"] == ref.result["n_used"]
    for k in ("median", "q1", "q3", "iqr", "lower_whisker", "upper_whisker"):
        assert r[k] == ref.result[k], k
    assert purity == ref.purity or abs(purity - ref.purity) <= 1e-6 * abs(ref.purity)
    assert purity == ref.purity, "restated in the same types: expected bit-identical"
    assert int(est.used.sum()) == r["n_used"]


@pytest.mark.skipif(not po.tap_available(), reason="reference tap not built")
@pytest.mark.parametrize("name", list(PURITY_CASES))
def test_purity_matches_reference(name):
    un, ut = pair(name)
    tp = somatic_cases.param_sets()["purity_q20"]           # estimate_purity defaults: -q 20, supplementary tagged
    ref = po.ReferencePurity(un, ut, tp)
    on, ot = po.OracleSomatic(un, tp, "extract_normal"), po.OracleSomatic(ut, tp, "extract_tumor")
    est = host.TumorPurityEstimator([as_result(on)], [as_result(ot)])
    purity = est.estimateTumorPurity()
    compare(est, purity, ref)
    if name == "no_support":
        assert purity == 0.0
    else:
        assert ref.result["n_used"] > 20 and 0.0 < purity <= 1.0, (ref.result, purity)


def test_purity_failure_modes():
    """No position passes the filters -> the reference reports 0.0 (TumorPurityEstimator.cpp:78-82)."""
    z = dict(ratios_d=np.zeros((5, 4)), read_hp_count=np.zeros((5, 9), np.int32))
    est = host.TumorPurityEstimator([z], [z])
    assert est.estimateTumorPurity() == 0.0 and est.result["ok"] == 0 and est.result["filtered_normal_imbalance_zero"] == 5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["purity_60"])
def test_gpu_purity_matches_reference(name):
    un, ut = pair(name)
    tp = somatic_cases.param_sets()["purity_q20"]
    ctx = host.Context(0)
    rn = host.ExtractNorDataChrProcessor(ctx, un, tp).processSingleChrom(un)
    rt = host.ExtractTumDataChrProcessor(ctx, ut, tp).processSingleChrom(ut)
    est = host.TumorPurityEstimator([rn], [rt])
    purity = est.estimateTumorPurity()
    on, ot = po.OracleSomatic(un, tp, "extract_normal"), po.OracleSomatic(ut, tp, "extract_tumor")
    est_o = host.TumorPurityEstimator([as_result(on)], [as_result(ot)])
    assert purity == est_o.estimateTumorPurity() and est.result == est_o.result
    if po.tap_available():
        compare(est, purity, po.ReferencePurity(un, ut, tp))
    ctx.close()
