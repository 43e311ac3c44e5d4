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


# ---- CIGAR ops of 4095 bases and more (escaped in the 16-bit stream the kernels read) and reads of several super-chunks ----
def long_op_contig(seed=7):
    rng = np.random.default_rng(seed)
    L = 60_000
    ref = "".join(rng.choice(list("ACGT"), L))
    # homopolymer runs so that the D-op rule has something to look at
    ref = list(ref)
    for s in range(500, L - 20, 997):
        ref[s:s + 5] = ref[s] * 5
    ref = "".join(ref)
    vs = []
    for p in range(200, L - 200, 173):
        alt = "ACGT"[("ACGT".index(ref[p]) + 1 + (p % 3)) % 4]
        if p % 11 == 0:
            vs.append((p, ref[p], ref[p] + "GG", p % 2))                  # insertion
        elif p % 13 == 0:
            vs.append((p, ref[p:p + 3], ref[p], p % 2))                   # deletion
        else:
            vs.append((p, ref[p], alt, p % 2))
    R = handmade.read_from_ref
    reads = [
        R(ref, 10, "4094M", "l00"),                       # largest length that fits the 12-bit field
        R(ref, 11, "4095M", "l01"),                       # smallest escaped length
        R(ref, 12, "4096M3I5000M2D3000M", "l02"),
        R(ref, 13, "100M5000N3000M", "l03"),              # escaped N op: variants under it are skipped
        R(ref, 14, "7S4200=1X800=9S", "l04"),             # = / X ops, counted clips around an escaped op
        R(ref, 15, "300M6000D2000M", "l05"),              # escaped D op over many variants (D-op rule on the first only)
        R(ref, 16, "5000M", "l06", edits={k: "A" for k in range(100, 4900, 37)}),
        R(ref, 20000, "4500M1I4500M", "l07"),
        R(ref, 30000, "12000M", "l08"),
    ]
    # reads of > 1536 and > 3072 ops (several super-chunks), with an escaped op in the middle of a later super-chunk
    def chopped(pos, n_units, name, long_at=None):
        cig = []
        for u in range(n_units):
            if long_at is not None and u == long_at:
                cig.append("4300M")
            cig.append("5M1I" if u % 3 else "4M1D")
        return R(ref, pos, "".join(cig) + "20M", name)
    reads += [chopped(100, 900, "m00"), chopped(101, 1700, "m01", long_at=1200), chopped(5000, 2500, "m02", long_at=50),
              chopped(9000, 800, "m03"), chopped(9001, 770, "m04")]
    # short reads in between shift the CIGAR offsets of their neighbours through every alignment (mod 8 ops)
    for k in range(23):
        reads.append(R(ref, 50 + 97 * k, "%dM" % (300 + k) if k % 4 else "%dM1I%dM" % (100 + k, 150), "s%02d" % k))
    reads.sort(key=lambda r: r["pos"])
    return handmade.ManualContig(ref, vs, reads)


def test_oracle_matches_reference_on_long_ops():
    po = pytest.importorskip("oracle.pyoracle")
    if not po.tap_available():
        pytest.skip("reference tap not built")
    c = long_op_contig()
    assert ((c.cigar >> 4) >= 4095).sum() >= 10 and c.n_cigar.max() > 3072
    p = ffi.default_phase_params(True)
    ref = po.ReferencePhase(c, p, stop_after_calls=True)
    a = po.OraclePhase(c, p, apply_filter=False, stages=1)
    cmp.assert_same_calls(cmp.calls_by_read(a.call_off, a.calls, c.var_pos), cmp.tap_stage_by_read(ref.stage_a), "long ops A")
    for k in ("clip_pos", "clip_front", "clip_back"):
        assert np.array_equal(getattr(a, k), getattr(ref, k)), k
    assert len(a.calls) > 500


@pytest.mark.gpu
def test_gpu_long_ops_match_oracle():
    """Escaped lengths, multi-super-chunk reads and every CIGAR offset alignment through the bulk-copy pipeline of k_call_alleles,
    in the phase dialect (uint32 and 16-bit submission) and the germline tag dialect."""
    po = pytest.importorskip("oracle.pyoracle")
    from .test_tag import check_gpu_tag
    c = long_op_contig()
    for is_ont in (True, False):
        parity.check_phase(c, ffi.default_phase_params(is_ont))
    p = ffi.default_phase_params(True)
    orc = po.OraclePhase(c, p)
    ps = np.where(orc.ps != 0, orc.ps, 1 + (np.arange(c.n_var) // 40) * 1000).astype(np.int32)    # every variant in some phase set
    hap = np.where(orc.ps != 0, orc.hap_ref == 1, (np.arange(c.n_var) % 2) == 1)
    ctx = host.Context(0)
    check_gpu_tag(c.phased(ps, hap), ffi.default_tag_params(), ctx)
    ctx.close()
