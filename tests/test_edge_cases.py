"""Edge cases of the allele-calling rules on hand-written contigs.  The CPU half (oracle vs live reference tap)
runs here; the -m gpu half runs the same contigs through the CUDA path."""
import importlib

import numpy as np
import pytest

from . import handmade, parity
from . import compare as cmp

ffi = importlib.import_module("longphase_s_b200._ffi")
host = importlib.import_module("longphase_s_b200.host")

# reference with a homopolymer (AAAAAA at 20..25) and a CA tandem repeat (CACACACACACA at 41..52)
REF = ("GCTAGCTAGGCTTACGGATC" "AAAAAA" "TGCCGTAGCTTGACC" "CACACACACACA" "GTTGCATGCATCGGATCGTTAGCTAGCATCGAT"
       "CGGATATCGCGATTAGCGCTAGCTAGGATCGATCGGCTAGCTAGCTAGGCTAGCTTAGGCTAGGATCGATTCGAGGCTAGCTAGGCTAGCTAGGCATTGCAG")


def variants():
    return [
        (5, "C", "T", 0),          # plain SNP
        (22, "A", "G", 0),         # SNP inside the homopolymer
        (23, "A", "C", 1),         # second homopolymer SNP 1 bp away -> filterSNP erases it on ONT
        (30, "T", "TGG", 0),       # insertion
        (40, "C", "CCA", 1),       # insertion in front of the CA repeat -> danger
        (60, "G", "A", 0),
        (70, "CGG", "C", 0),       # deletion
        (90, "AT", "GC", 0),       # complex (MNP): never called
        (100, "G", "T", 1),
        (130, "T", "C", 0),
        (150, "A", "G", 1),
    ]


def reads():
    R = handmade.read_from_ref
    out = [
        R(REF, 0, "120M", "r00"),                                              # all REF
        R(REF, 0, "120M", "r01", edits={5: "T", 22: "G", 60: "A", 100: "T"}),   # ALT at SNPs
        R(REF, 0, "120M", "r02", edits={5: "G"}),                               # third base: no call at 5
        R(REF, 0, "31M2I89M", "r03"),                                          # insertion variant called ALT (M ends at 30)
        R(REF, 0, "30M2I90M", "r04"),                                          # variant lands in the last op: no call
        R(REF, 0, "29M2I10M1D80M", "r04b"),                                    # insertion one base early: REF allele
        R(REF, 0, "41M2I79M", "r05"),                                          # danger insertion ALT (quality -5)
        R(REF, 0, "71M2D47M", "r06"),                                          # deletion variant ALT
        R(REF, 0, "71M3D46M", "r07"),                                          # deletion of another length still counts
        R(REF, 0, "21M3D96M", "r08"),                                          # D op over the homopolymer SNPs: D-op rule on the first only
        R(REF, 0, "58M5D57M", "r09"),                                          # D op over a non-homopolymer SNP: nothing
        R(REF, 0, "69M4D47M", "r10"),                                          # D op covering the deletion variant itself (pos 70 not homopolymer)
        R(REF, 0, "10S120M8S", "r11"),                                         # counted clips (front, back)
        R(REF, 0, "5S120M3S", "r12"),                                          # short clips: not counted
        R(REF, 2, "6H118M7H", "r13"),                                          # hard clips counted, front iff index 0
        R(REF, 2, "50M20N48M", "r14"),                                         # N op skips variant 60
        R(REF, 2, "10=1X107=", "r15"),                                         # = and X ops
        R(REF, 2, "30M2P88M", "r16"),                                          # P op
        R(REF, 3, "117M", "r17", mapq=0),                                      # filtered by MAPQ
        R(REF, 3, "117M", "r18", flag=0x100),                                  # secondary: filtered
        R(REF, 3, "117M", "r19", flag=0x400),                                  # duplicate: filtered
        R(REF, 3, "117M", "r20", flag=0x800),                                  # supplementary: kept
        R(REF, 4, "116M", "r21", flag=0x4),                                    # unmapped flag: filtered
        R(REF, 95, "30M", "r22"),                                              # short read with few variants
        R(REF, 125, "40M", "r23"),
        dict(name="r24", pos=4, cigar="116M", seq="", qual=[]),                 # SEQ '*': aborted at the first variant
        dict(name="r25", pos=4, cigar="7S116M", seq="", qual=[]),               # clip before the abort stays counted
        R(REF, 6, "24M1I90M", "r26"),                                          # insertion variant at the LAST op boundary handled below
        R(REF, 126, "25M", "r27"),                                             # covers only variant 130 and 150
        R(REF, 151, "20M", "r28"),                                             # starts after the last variant: outside the iterator region
        R(REF, 150, "20M", "r29"),                                             # starts AT the last variant: outside "chr:1-lastSNP"
    ]
    # several reads sharing the low-quality pattern so that edges get 0.1 contributions
    for k in range(6):
        out.append(R(REF, 1, "119M", f"s{k:02d}", qual=5 if k % 2 else 30, edits={5: "T"} if k % 3 == 0 else None))
    out.sort(key=lambda r: r["pos"])
    return out


def contig():
    return handmade.ManualContig(REF, variants(), reads())


def test_oracle_matches_reference_on_edge_cases():
    po = pytest.importorskip("oracle.pyoracle")
    if not po.tap_available():
        pytest.skip("reference tap not built")
    c = contig()
    for is_ont in (True, False):
        p = ffi.default_phase_params(is_ont)
        ref = po.ReferencePhase(c, p, stop_after_calls=True)
        a = po.OraclePhase(c, p, apply_filter=False, stages=1)
        b = po.OraclePhase(c, p, apply_filter=True, stages=1)
        cmp.assert_same_calls(cmp.calls_by_read(a.call_off, a.calls, c.var_pos), cmp.tap_stage_by_read(ref.stage_a), "edge A")
        cmp.assert_same_calls(cmp.calls_by_read(b.call_off, b.calls, c.var_pos), cmp.tap_stage_by_read(ref.stage_b), "edge B", allow_empty_in_b=True)
        for k in ("clip_pos", "clip_front", "clip_back"):
            assert np.array_equal(getattr(a, k), getattr(ref, k)), k
        # the rules we meant to hit did fire
        by = cmp.calls_by_read(a.call_off, a.calls, c.var_pos)
        names = [c.name(i) for i in range(c.n_reads)]
        idx = {n: i for i, n in enumerate(names)}
        assert list(by[idx["r03"]][1][by[idx["r03"]][0] == 30]) == [1]
        assert list(by[idx["r04"]][1][by[idx["r04"]][0] == 30]) == []      # variant in the LAST op: i+1 < n_cigar fails, no call
        assert list(by[idx["r00"]][1][by[idx["r00"]][0] == 30]) == []      # single-op read: same
        assert list(by[idx["r04b"]][1][by[idx["r04b"]][0] == 30]) == [0]
        assert -5 in by[idx["r05"]][2] and list(by[idx["r06"]][1][by[idx["r06"]][0] == 70]) == [1]
        assert (a.calls["origin"] == 1).sum() >= 1                      # D-op rule
        assert a.read_status[idx["r24"]] == 1 and a.read_status[idx["r25"]] == 1   # aborted
        assert a.read_status[idx["r28"]] == 2 and a.read_status[idx["r29"]] == 2   # outside the iterator region
        assert 90 not in np.concatenate([v[0] for v in by.values()])     # the MNP is never called


def test_unsupported_cigar_op_is_an_error_in_the_oracle():
    po = pytest.importorskip("oracle.pyoracle")
    c = handmade.ManualContig(REF, variants(), [dict(name="bad", pos=0, cigar=[(50, 0), (3, 9), (60, 0)], seq="A" * 110, qual=[30] * 110)])
    o = po.OraclePhase(c, ffi.default_phase_params(True), stages=1)
    assert o.rc == -4   # LPS_E_CIGAR; the reference prints and exit(1)s (ParsingBam.cpp:1625-1628)


@pytest.mark.gpu
def test_gpu_edge_cases_match_oracle():
    c = contig()
    for is_ont in (True, False):
        parity.check_phase(c, ffi.default_phase_params(is_ont))


@pytest.mark.gpu
def test_gpu_unsupported_cigar_op():
    c = handmade.ManualContig(REF, variants(), [dict(name="bad", pos=0, cigar=[(50, 0), (3, 9), (60, 0)], seq="A" * 110, qual=[30] * 110)])
    ctx = host.Context(0)
    p = ffi.default_phase_params(True)
    bp = host.BamParser(ctx, c, p)
    with pytest.raises(host.LpsError) as e:
        bp.direct_detect_alleles(c)
    assert e.value.code == -4
    ctx.close()


@pytest.mark.gpu
def test_gpu_empty_and_degenerate_batches():
    po = pytest.importorskip("oracle.pyoracle")
    ctx = host.Context(0)
    p = ffi.default_phase_params(True)
    # no reads at all
    c0 = handmade.ManualContig(REF, variants(), [])
    bp = host.BamParser(ctx, c0, p)
    r = bp.direct_detect_alleles(c0)
    assert r["n_calls"] == 0 and len(r["clip_pos"]) == 0
    g = host.VairiantGraph(ctx, p)
    e = g.addEdge()
    assert e["n_nodes"] == 0
    res = g.phasingProcess()
    assert (res["ps"] == 0).all()
    # reads that overlap no variant, a read with zero CIGAR ops, a single-variant read
    R = handmade.read_from_ref
    c1 = handmade.ManualContig(REF, variants(), [R(REF, 6, "10M", "a"), dict(name="b", pos=8, cigar=[], seq="", qual=[]),
                                                 R(REF, 55, "10M", "c"), R(REF, 56, "10M", "d")])
    parity.check_phase(c1, p, ctx=ctx)
    ctx.close()
