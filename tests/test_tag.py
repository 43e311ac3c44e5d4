"""Germline haplotag: oracle vs the live reference tap (CPU) and CUDA vs oracle (-m gpu)."""
import importlib

import numpy as np
import pytest

from . import compare as cmp
from . import tag_cases

po = pytest.importorskip("oracle.pyoracle")
host = importlib.import_module("longphase_s_b200.host")


@pytest.mark.skipif(not po.tap_available(), reason="reference tap not built")
@pytest.mark.parametrize("name", tag_cases.TAG_CASES)
@pytest.mark.parametrize("kind", ["phase_result", "blocks50"])
def test_tag_oracle_matches_reference(name, kind):
    c = tag_cases.get(name, kind)
    for pname, tp in tag_cases.param_sets().items():
        ref = po.ReferenceTag(c, tp)
        orc = po.OracleTag(c, tp)
        assert ref.rc == 0 and orc.rc == 0
        n_ok = cmp.assert_tag_matches_reference(orc, ref, c)
        assert n_ok > 0.7 * c.n_reads
        # ReadStatistics derived from the oracle's per-read outputs equal the reference's counters
        if cmp.well_formed_reads(c).all():
            proc = orc.category == 0
            assert ref.stats["total_tag"] == int((orc.hp[proc] != 0).sum()), pname
            assert ref.stats["total_hp1"] == int((orc.hp == 1).sum()) and ref.stats["total_hp2"] == int((orc.hp == 2).sum())
            assert ref.stats["total_alignment"] == c.n_reads


def check_gpu_tag(c, tp, ctx):
    orc = po.OracleTag(c, tp)
    proc = host.GermlineHaplotagChrProcessor(ctx, c, tp)
    res = proc.processSingleChrom(c)
    assert np.array_equal(res["category"], orc.category), "dispatch categories differ"
    for k in ("hp", "ps", "pq", "h1", "h2"):
        assert np.array_equal(res[k], getattr(orc, k)), f"{k} differs"
    assert np.array_equal(res["call_off"], orc.call_off) and res["calls"].tobytes() == orc.calls.tobytes(), "per-read variant lists differ"
    st, p = res["stats"], orc.category == 0
    assert st["total_alignment"] == c.n_reads and st["total_tag"] == int((orc.hp != 0).sum())
    assert st["total_hp1"] == int((orc.hp == 1).sum()) and st["total_hp2"] == int((orc.hp == 2).sum())
    assert st["total_lower_quality"] == int((orc.category == 1).sum()) and st["total_secondary"] == int((orc.category == 3).sum())
    assert st["total_without_variant"] == int(((orc.h1 == 0) & (orc.h2 == 0) & p).sum())
    res2 = ctx.tag_reads(tp, want_calls=False)
    for k in ("hp", "ps", "pq", "h1", "h2", "category"):
        assert np.array_equal(res2[k], res[k])
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("name", tag_cases.TAG_CASES)
def test_gpu_tag_matches_oracle(name):
    ctx = host.Context(0)
    tagged = 0
    for kind in ("phase_result", "blocks50"):
        c = tag_cases.get(name, kind)
        for tp in tag_cases.param_sets().values():
            res = check_gpu_tag(c, tp, ctx)
            tagged += int((res["hp"] != 0).sum())
    assert tagged > 0
    ctx.close()


@pytest.mark.gpu
def test_gpu_tag_requires_phased_variant_table():
    from . import cases
    c, _ = cases.get("snp_only")
    ctx = host.Context(0)
    ctx.set_reference(c.ref)
    ctx.set_variants(c.variants_struct(), 0)          # no PS column
    ctx.submit(c.batch_struct())
    with pytest.raises(host.LpsError):
        ctx.tag_reads(tag_cases.param_sets()["default"])
    ctx.close()


@pytest.mark.gpu
def test_gpu_tag_rejects_interleaved_rows():
    """lps_read_batch.sq serves the phase calls only: the tag dialects read runs of bases and say so instead of reading garbage."""
    name = next(iter(tag_cases.TAG_CASES))
    c = tag_cases.get(name, "phase_result")
    tp = tag_cases.param_sets()["default"]
    ctx = host.Context(0)
    check_gpu_tag(c, tp, ctx)                          # context holds the tagged variant table now
    ctx.submit(c.batch_struct_sq())
    with pytest.raises(host.LpsError, match="phase calls only"):
        ctx.tag_reads(tp)
    ctx.submit(c.batch_struct())
    assert np.array_equal(ctx.tag_reads(tp, want_calls=False)["hp"], po.OracleTag(c, tp).hp)
    ctx.close()
