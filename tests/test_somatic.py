"""Somatic family (extract-normal, extract-tumor incl. window diff, somatic tagging): oracle vs the live reference tap (CPU)
and CUDA vs oracle (-m gpu)."""
import importlib

import numpy as np
import pytest

from . import somatic_cases

po = pytest.importorskip("oracle.pyoracle")
host = importlib.import_module("longphase_s_b200.host")

MODES = ["extract_normal", "extract_tumor", "somatic_tag"]
PER_SLOT = ["pos_base", "read_hp_count", "somatic_read_hp_count", "case_count", "window_hist", "hp_before_count", "hp_after_count",
            "h3_before_count", "h3_after_count", "cover_start", "cover_end", "ratios_f", "ratios_d", "case_read_count"]
# SomaticData::alleleCount is never initialised by the reference (HaplotagType.h:284-293: absent from the constructor's
# initialiser list) and never read; it is compared between the CUDA path and the oracle only.


def well_formed(c):
    """Alignments whose SEQ is present: a `*` SEQ makes the reference read past the record (undefined)."""
    return c.l_qseq > 0


def pick(c, mode):
    un, ut = c
    return un if mode == "extract_normal" else ut


@pytest.mark.skipif(not po.tap_available(), reason="reference tap not built")
@pytest.mark.parametrize("name", list(somatic_cases.SOMATIC_CASES))
@pytest.mark.parametrize("mode", MODES)
def test_somatic_oracle_matches_reference(name, mode):
    c = pick(somatic_cases.get(name), mode)
    assert well_formed(c).all()
    for pname, tp in somatic_cases.param_sets().items():
        orc, ref = po.OracleSomatic(c, tp, mode), po.ReferenceSomatic(c, tp, mode)
        assert orc.rc == 0 and ref.rc == 0
        assert np.array_equal(orc.tum_var, ref.tum_var) and np.array_equal(orc.category, ref.category), pname
        for k in PER_SLOT:
            assert np.array_equal(getattr(orc, k), getattr(ref, k)), (pname, k)
        proc = orc.category == 0
        if mode == "extract_normal":
            for k in ("read_hp", "ps", "pq", "h1", "h2", "n_ps", "end_pos", "read_len"):
                assert np.array_equal(getattr(orc, k)[proc], getattr(ref, k)[proc]), (pname, k)
            assert orc.pos_base.sum() > 0 and orc.read_hp_count[:, 1:3].sum() > 0
        elif mode == "extract_tumor":
            # the reference keeps a per-read record only for alignments that cover a tumor position with MAPQ >= q
            seen = proc & (ref.read_hp != -1)
            has_rec = np.zeros(c.n_reads, bool)
            for r in np.nonzero(proc)[0]:
                cs = orc.calls[int(orc.call_off[r]):int(orc.call_off[r + 1])]
                has_rec[r] = bool((cs["quality"] & 1).any())
            assert np.array_equal(seen, has_rec), pname
            for k in ("read_hp", "ps", "h1", "h2", "h3", "n_ps", "end_pos", "read_len"):
                assert np.array_equal(getattr(orc, k)[seen], getattr(ref, k)[seen]), (pname, k)
            n_full = 0
            for r in np.nonzero(seen)[0]:
                a = orc.calls[int(orc.call_off[r]):int(orc.call_off[r + 1])]
                b = ref.calls[int(ref.call_off[r]):int(ref.call_off[r + 1])]
                if ref.pq[r] == 1:      # posHpPairs was recorded: the whole variantsHP map of the read is observable
                    assert np.array_equal(a["var"], b["var"]) and np.array_equal(a["allele"], b["allele"]), (pname, r)
                    assert np.array_equal(a["quality"] & 1, b["quality"] & 1)
                    n_full += 1
                else:                   # only tumorPosReadCorrBaseHP is
                    a = a[(a["quality"] & 1) != 0]
                    assert np.array_equal(a["var"], b["var"]) and np.array_equal(a["allele"], b["allele"]), (pname, r)
            assert n_full > 0 and orc.window_hist.sum() > 0 and orc.case_count.sum() > 0
            assert orc.n_window_items == int(orc.allele_count.sum())
        else:
            for k in ("read_hp", "ps", "pq"):
                assert np.array_equal(getattr(orc, k)[proc], getattr(ref, k)[proc]), (pname, k)
            st = stats_from(orc, c, tp)
            assert st == ref.stats, pname
            assert orc.hp_after_count.sum() > 0


def stats_from(o, c, tp):
    """ReadStatistics (HaplotagProcess.h:21-45) recomputed from the per-alignment products, as lps_somatic_tag_reads does."""
    cat, proc = o.category, o.category == 0
    mx, mn = np.maximum(o.h1, o.h2).astype(float), np.minimum(o.h1, o.h2).astype(float)
    with np.errstate(invalid="ignore", divide="ignore"):
        nsim = np.where(mx == 0, 0.0, mx / (mx + mn))
    pct = tp.percentage_threshold
    high = proc & np.where(o.h3 != 0, not (1.0 >= pct), (mx != 0) & ~(nsim >= pct))
    st = dict(total_alignment=c.n_reads, total_supplementary=int((cat == 4).sum() + (proc & ((c.flag & 0x800) != 0)).sum()),
              total_secondary=int((cat == 3).sum()), total_unmapped=int((cat == 2).sum()), total_tag=int((proc & (o.read_hp != 0)).sum()),
              total_untag=int((~proc).sum() + (proc & (o.read_hp == 0)).sum()), total_lower_quality=int((cat == 1).sum()),
              total_other_case=int((cat == 6).sum()), total_empty_variant=int((cat == 5).sum()), total_high_similarity=int(high.sum()),
              total_cross_two_block=int((proc & (o.n_ps > 1)).sum()), total_without_variant=int((proc & (mx == 0) & (o.h3 == 0)).sum()),
              total_read_only_h3=int((proc & (o.h1 == 0) & (o.h2 == 0) & (o.h3 != 0) & (o.read_hp == 3)).sum()))
    for k in range(9):
        st[f"hp{k}"] = int((proc & (o.read_hp == k)).sum())
    return st


def check_gpu_somatic(c, tp, mode, ctx):
    orc = po.OracleSomatic(c, tp, mode)
    cls = {"extract_normal": host.ExtractNorDataChrProcessor, "extract_tumor": host.ExtractTumDataChrProcessor,
           "somatic_tag": host.SomaticHaplotagChrProcessor}[mode]
    res = cls(ctx, c, tp).processSingleChrom(c)
    assert np.array_equal(res["tum_var"], orc.tum_var) and np.array_equal(res["category"], orc.category)
    for k in ("read_hp", "ps", "pq", "h1", "h2", "h3", "n_ps", "end_pos", "read_len"):
        assert np.array_equal(res[k], getattr(orc, k)), (mode, k)
    if mode == "somatic_tag":
        keys = ["hp_before_count", "hp_after_count", "h3_before_count", "h3_after_count", "cover_start", "cover_end", "hp_before"]
        assert res["derive_similarity"].tobytes() == orc.derive_similarity.tobytes()
        assert res["stats"] == stats_from(orc, c, tp)
    elif mode == "extract_tumor":
        keys = ["pos_base", "read_hp_count", "somatic_read_hp_count", "case_count", "allele_count", "window_hist", "ratios_f", "ratios_d",
                "case_read_count"]
        assert res["n_window_items"] == orc.n_window_items
    else:
        keys = ["pos_base", "read_hp_count", "ratios_f", "ratios_d"]
    for k in keys:
        assert np.array_equal(res[k], getattr(orc, k)), (mode, k)
    if mode != "extract_normal":
        assert np.array_equal(res["call_off"], orc.call_off) and res["calls"].tobytes() == orc.calls.tobytes(), "per-read variant lists differ"
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(somatic_cases.SOMATIC_CASES))
def test_gpu_somatic_matches_oracle(name):
    ctx = host.Context(0)
    pair = somatic_cases.get(name)
    for mode in MODES:
        for tp in somatic_cases.param_sets().values():
            check_gpu_somatic(pick(pair, mode), tp, mode, ctx)
    ctx.close()


@pytest.mark.gpu
def test_gpu_somatic_noseq_and_state_errors():
    """`*` SEQ alignments (hits beyond l_qseq are dropped, as in the oracle) and call-order errors."""
    import copy
    un, ut = somatic_cases.get("snv_indel")
    ctx = host.Context(0)
    ctx.set_reference(ut.ref)
    ctx.set_variants(ut.variants_struct(), 0)
    ctx.submit(ut.batch_struct())
    with pytest.raises(host.LpsError):
        ctx.extract_tumor(somatic_cases.param_sets()["purity_q20"])       # no tumor table yet
    bad = copy.copy(ut)
    bad.is_somatic = np.ones(ut.n_var, np.uint8)                         # somatic flag on positions without a TUMOR record
    with pytest.raises(host.LpsError):
        ctx.set_tumor_variants(bad.tumor_struct())
    c = copy.copy(ut)
    c.l_qseq = ut.l_qseq.copy()
    c.l_qseq[::7] = 0
    for mode in MODES:
        check_gpu_somatic(c, somatic_cases.param_sets()["purity_q20"], mode, ctx)
    ctx.close()
