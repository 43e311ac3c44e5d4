"""Edge cases of the tag dialects (germline haplotag and the somatic family) on a hand-written union map: N / P / = / X ops
(the window diff treats N, P and X as "consume the iteration, do not move"), hard and soft clips, D ops over NORMAL and TUMOR
variants, MNP records, NORMAL + TUMOR records at one position with different ALT bases, unphased NORMAL records, variants in the
first / last op, MAPQ and flag dispatch.  CPU: oracle vs the live reference tap; -m gpu: CUDA vs oracle."""
import importlib

import numpy as np
import pytest

from . import handmade
from .test_edge_cases import REF, reads
from .test_somatic import MODES, PER_SLOT, check_gpu_somatic

ffi = importlib.import_module("longphase_s_b200._ffi")
host = importlib.import_module("longphase_s_b200.host")


def entries():
    return [
        dict(pos=5, nor=("C", "T", 0, 1)),                                         # phased germline SNP, block 1
        dict(pos=12, tum=("T", "G", 2), somatic=1, derive=1),                      # somatic SNV, unphased het, derives from H1
        dict(pos=22, nor=("A", "G", 0, 1), tum=("A", "C", 2)),                     # NORMAL and TUMOR records with different ALT, homopolymer
        dict(pos=23, tum=("A", "C", 3), somatic=1, derive=2),                      # tumor homozygous SNV inside the homopolymer
        dict(pos=30, nor=("T", "TGG", 0, 1)),                                      # germline insertion
        dict(pos=36, tum=("T", "TAA", 1, 37), somatic=1),                          # phased tumor insertion with its own phase set
        dict(pos=40, nor=("C", "CCA", 1, 1)),                                      # germline insertion, HP1 carries ALT
        dict(pos=47, tum=("CAC", "C", 2), somatic=1, derive=1),                    # somatic deletion
        dict(pos=60, nor=("G", "A", 0, 61)),                                       # second phase set: reads spanning both are untagged
        dict(pos=70, nor=("CGG", "C", 0, 61), tum=("CGG", "C", 2)),                # germline deletion also reported in the tumor VCF
        dict(pos=90, nor=("AT", "GC", 0, 61)),                                     # germline MNP: never votes
        dict(pos=95, tum=("GC", "AT", 2), somatic=1),                              # tumor MNP: positions only
        dict(pos=100, nor=("G", "T", 1, 61, 2)),                                   # NORMAL record that is not PHASED_HETERO
        dict(pos=110, tum=("A", "G", 2)),                                          # tumor-only germline leak: not a somatic variant
        dict(pos=130, nor=("T", "C", 0, 61)),
        dict(pos=150, tum=("A", "G", 2), somatic=1, derive=0),
    ]


def extra_reads():
    R = handmade.read_from_ref
    return [R(REF, 0, "45M5D70M", "t00"),                                           # D op over the somatic deletion (and nothing else)
            R(REF, 0, "37M2I83M", "t01"),                                           # phased tumor insertion ALT
            R(REF, 0, "48M2D70M", "t02", edits={12: "G"}),                          # somatic deletion ALT + somatic SNV ALT
            R(REF, 0, "186M", "t03", edits={12: "G", 23: "C", 150: "G", 5: "T"}),   # several somatic ALTs on an H2 read
            R(REF, 0, "180M4S", "t04", edits={12: "G", 23: "C", 150: "G"}),           # ... on an H1 read
            R(REF, 0, "120M", "t05", mapq=5, edits={12: "G"}),                      # low MAPQ: counted in PosBase only
            R(REF, 1, "11=1X2=1X100=", "t06", edits={11: "G", 14: "T"}),            # somatic SNV inside an X op: the window scan never moves
            R(REF, 1, "60M30N50M", "t07"),                                          # N op over the MNPs
            R(REF, 8, "3S10M3I20M2D60M4S", "t08", edits={7: "G"}),                  # tumor SNV in the first M op after a soft clip
            R(REF, 100, "60M", "t09", edits={50: "G"}),                             # variant 150 near the end of the read
            R(REF, 149, "2M5I35M", "t10", edits={1: "G"})]                          # variant in a 2-base first op


def union(with_noseq):
    rs = [r for r in reads() if with_noseq or len(r["seq"]) > 0] + extra_reads()
    rs.sort(key=lambda r: r["pos"])
    # the reference string must cover the reads; REF ends at 223
    return handmade.ManualUnion(REF, entries(), rs)


def params():
    T = ffi.LpsTagParams
    return [T(mapping_quality=1, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6),
            T(mapping_quality=20, mapq_filter=1, tag_supplementary=0, have_reference=1, percentage_threshold=0.6),
            T(mapping_quality=1, mapq_filter=1, tag_supplementary=1, have_reference=0, percentage_threshold=0.9)]


def test_somatic_edge_cases_oracle_matches_reference():
    po = pytest.importorskip("oracle.pyoracle")
    if not po.tap_available():
        pytest.skip("reference tap not built")
    c = union(with_noseq=False)
    for tp in params():
        for mode in MODES:
            orc, ref = po.OracleSomatic(c, tp, mode), po.ReferenceSomatic(c, tp, mode)
            assert orc.rc == 0 and ref.rc == 0
            assert np.array_equal(orc.category, ref.category), mode
            for k in PER_SLOT:
                assert np.array_equal(getattr(orc, k), getattr(ref, k)), (mode, k)
            proc = orc.category == 0
            if mode == "extract_tumor":
                proc = proc & (ref.read_hp != -1)
                keys = ("read_hp", "h1", "h2", "h3", "n_ps", "end_pos", "read_len")
            elif mode == "somatic_tag":
                keys = ("read_hp", "ps", "pq")
            else:
                keys = ("read_hp", "ps", "pq", "h1", "h2", "n_ps", "end_pos", "read_len")
            for k in keys:
                assert np.array_equal(getattr(orc, k)[proc], getattr(ref, k)[proc]), (mode, k)
    # the cases we meant to hit did fire (default parameters, tumor pass)
    o = po.OracleSomatic(c, params()[0], "extract_tumor")
    slot = {int(c.var_pos[v]): i for i, v in enumerate(o.tum_var)}
    assert o.window_hist.sum() > 0 and o.pos_base[slot[47], ffi.PB_FIELDS.index("del")] > 0      # read D op over the somatic deletion
    assert o.allele_count[slot[47], 1] > 0 and o.allele_count[slot[36], 1] > 0 and o.allele_count[slot[12], 1] >= 3
    assert o.read_hp_count[slot[95]].sum() > 0 and o.pos_base[slot[95]].sum() == 0                  # tumor MNP: read HP only
    assert o.case_count.sum() > 0


def test_germline_tag_edge_cases_oracle_matches_reference():
    po = pytest.importorskip("oracle.pyoracle")
    if not po.tap_available():
        pytest.skip("reference tap not built")
    from . import compare as cmp
    full = union(with_noseq=False)
    keep = np.nonzero((full.nor_present != 0) & (full.var_gt_kind == 1))[0]       # what VcfParser stores for `haplotag`
    c = handmade.ManualContig(REF, [(int(full.var_pos[i]),) + full.variant_strings(i) + (int(full.var_hp1_is_alt[i]),) for i in keep],
                              sorted([r for r in reads() if len(r["seq"]) > 0] + extra_reads(), key=lambda r: r["pos"]))
    c.var_ps = np.ascontiguousarray(full.var_ps[keep])
    c.var_gt_kind = np.ones(len(keep), np.uint8)
    for tp in params():
        ref, orc = po.ReferenceTag(c, tp), po.OracleTag(c, tp)
        assert cmp.assert_tag_matches_reference(orc, ref, c) > 0


@pytest.mark.gpu
def test_gpu_tag_edge_cases_match_oracle():
    ctx = host.Context(0)
    c = union(with_noseq=True)
    for tp in params():
        for mode in MODES:
            check_gpu_somatic(c, tp, mode, ctx)
    ctx.close()
