"""Parity at bench-like sizes (-m gpu): one 16 Mb / 30x contig (25 k reads, 1330 CIGAR ops per read — the shape the bench runs) through
every dialect against the oracle, a C5-like dense contig (120x, 50 kb reads, 1 variant / 300 bp: candidate-buffer overflow pass,
multi-super-chunk reads), and size-independent properties of the 64 Mb bench contig itself (pinned-host and device-resident
submission give identical bytes; repeated calls are idempotent; CSR offsets are sorted and consistent)."""
import importlib

import numpy as np
import pytest

from . import cases, parity  # noqa: F401
from .test_tag import check_gpu_tag
from .test_somatic import MODES, check_gpu_somatic

synth = importlib.import_module("longphase_s_b200.synth")
ffi = importlib.import_module("longphase_s_b200._ffi")
host = importlib.import_module("longphase_s_b200.host")
workloads = importlib.import_module("longphase_s_b200.workloads")


@pytest.mark.gpu
def test_gpu_bench_shape_contig_matches_oracle():
    po = pytest.importorskip("oracle.pyoracle")
    c = synth.Contig(seed=41, contig_len=16_000_000, indel_frac=0.1, depth=30.0, mean_len=20000.0)
    p = ffi.default_phase_params(True)
    info = parity.check_phase(c, p)
    assert info["reads"] > 20_000 and info["calls"] > 400_000
    orc = po.OraclePhase(c, p)
    ctx = host.Context(0)
    check_gpu_tag(c.phased(orc.ps, orc.hap_ref == 1), ffi.default_tag_params(), ctx)
    ctx.close()


@pytest.mark.gpu
def test_gpu_timed_64mb_contig_matches_oracle_at_every_stage():
    """The contig bench.py times first (workloads.weak_seed(0, 0), 64 Mb, 30x, ~99 k reads, 1.8 M calls): every stage of the phase path
    and the germline tagging pass against the oracle, at the timed size; its digest must be the committed one bench.py gates on."""
    po = pytest.importorskip("oracle.pyoracle")
    c = synth.Contig(**workloads.phase_kwargs(workloads.weak_seed(0, 0)))
    p = ffi.default_phase_params(True)
    info = parity.check_phase(c, p)
    assert info["reads"] > 90_000 and info["calls"] > 1_500_000
    orc = po.OraclePhase(c, p)
    want = workloads.load_digests().get(workloads.key_of(workloads.phase_kwargs(workloads.weak_seed(0, 0))))
    assert want is not None and want["digest"] == parity.oracle_phase_digest(orc, c.n_reads), "committed bench digest is stale"
    ctx = host.Context(0)
    check_gpu_tag(c.phased(orc.ps, orc.hap_ref == 1), ffi.default_tag_params(), ctx)
    ctx.close()


@pytest.mark.gpu
def test_gpu_timed_c4_shard_matches_oracle():
    """The C4 shard bench.py times (tumor 50x / normal 25x pair over one 32 Mb contig): the three somatic passes against the oracle."""
    un, ut = workloads.c4_pair(synth, 32.0)
    tp = ffi.LpsTagParams(mapping_quality=20, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6)
    ctx = host.Context(0)
    for mode in MODES:
        res = check_gpu_somatic(un if mode == "extract_normal" else ut, tp, mode, ctx)
        assert res["n_tum"] > 1000
    ctx.close()


@pytest.mark.gpu
def test_gpu_dense_stress_contig_matches_oracle():
    kw = dict(seed=42, contig_len=1_500_000, indel_frac=0.1, depth=120.0, mean_len=50000.0, variant_rate=1 / 300.0, somatic_rate=1 / 3000.0)
    c = synth.Contig(**kw)
    info = parity.check_phase(c, ffi.default_phase_params(True))
    assert info["calls"] / info["reads"] > 100                 # ~166 calls per read: reads beyond the shared candidate buffer exist
    cn = synth.Contig(**kw, purity=0.0, read_seed=421)
    un = cn.somatic_union(seed=4)
    ut = un.with_reads_of(synth.Contig(**kw, purity=0.5, read_seed=422))
    ctx = host.Context(0)
    tp = ffi.LpsTagParams(mapping_quality=20, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6)
    for mode in MODES:
        check_gpu_somatic(un if mode == "extract_normal" else ut, tp, mode, ctx)
    ctx.close()


@pytest.mark.gpu
def test_gpu_full_size_properties():
    import ctypes as C
    import torch
    c = synth.Contig(**workloads.phase_kwargs(workloads.weak_seed(0, 0)))     # the bench's first contig
    p = ffi.default_phase_params(True)
    ctx = host.Context(0)
    ctx.set_reference(c.ref)
    vs = c.variants_struct()
    ctx.set_variants(vs, True)
    ctx.submit(c.batch_struct())                                 # pageable host buffers: everything is copied
    a = ctx.call_alleles(p, want_host=True)
    r_a = ctx.phase_contig(p)
    # device-resident submission of the same batch
    names = ["ref_start", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "flag", "mapq", "name_rank", "cigar", "seq4", "qual"]
    ptypes = dict(ref_start=ffi.i32p, l_qseq=ffi.i32p, n_cigar=ffi.u32p, cigar_off=ffi.u64p, seq_off=ffi.u64p, qual_off=ffi.u64p,
                  flag=ffi.u16p, mapq=ffi.u8p, name_rank=ffi.i32p, cigar=ffi.u32p, seq4=ffi.u8p, qual=ffi.u8p)
    dt = {k: torch.from_numpy(getattr(c, k).view(np.uint8).reshape(-1)).cuda() for k in names}
    pt = {k: torch.from_numpy(getattr(c, k).view(np.uint8).reshape(-1)).pin_memory() for k in names}

    def batch(d):
        return ffi.LpsReadBatch(n_reads=c.n_reads, cigar_len=len(c.cigar), seq_bytes=len(c.seq4), qual_bytes=len(c.qual),
                                **{k: C.cast(d[k].data_ptr(), ptypes[k]) for k in names})
    ctx.submit_device(batch(dt))
    b = ctx.call_alleles(p, want_host=True)
    r_b = ctx.phase_contig(p)
    ctx.submit(batch(pt))                                        # pinned host buffers: SEQ / QUAL gathered over PCIe (zero-copy)
    cc = ctx.call_alleles(p, want_host=True)
    r_c = ctx.phase_contig(p)
    # SEQ + QUAL as interleaved rows (lps_read_batch.sq): pinned rows gathered over PCIe, then the same rows resident on the device
    sq, sq_off = c.pack_sq()
    sq_p, off_p = torch.from_numpy(sq).pin_memory(), torch.from_numpy(sq_off.view(np.int64)).pin_memory()
    sq_d, off_d = torch.from_numpy(sq).cuda(), torch.from_numpy(sq_off.view(np.int64)).cuda()

    def batch_sq(d, sq_t, off_t):
        bb = batch(d)
        bb.seq4, bb.qual, bb.qual_off = C.cast(None, ffi.u8p), C.cast(None, ffi.u8p), C.cast(None, ffi.u64p)
        bb.seq_bytes = bb.qual_bytes = 0
        bb.seq_off = C.cast(off_t.data_ptr(), ffi.u64p)
        bb.sq, bb.sq_bytes = C.cast(sq_t.data_ptr(), ffi.u8p), sq_t.numel()
        return bb
    s0 = ctx.stats()["h2d_bytes"]
    ctx.submit(batch_sq(pt, sq_p, off_p))
    dd = ctx.call_alleles(p, want_host=True)
    r_d = ctx.phase_contig(p)
    moved = ctx.stats()["h2d_bytes"] - s0
    assert moved < len(c.cigar) * 4 + 64 * c.n_reads + 2 * 40 * a["n_calls"] + (1 << 20), "pinned sq rows were copied instead of gathered"
    ctx.submit_device(batch_sq(dt, sq_d, off_d))
    ee = ctx.call_alleles(p, want_host=True)
    r_e = ctx.phase_contig(p)
    ctx.submit_device(batch(dt))
    for x in (b, cc, dd, ee):
        assert np.array_equal(a["call_off"], x["call_off"]) and a["calls"].tobytes() == x["calls"].tobytes()
        assert np.array_equal(a["read_status"], x["read_status"]) and np.array_equal(a["clip_pos"], x["clip_pos"])
    for x in (r_b, r_c, r_d, r_e):
        for k in ("ps", "hap_ref", "read_hp", "hp_counts"):
            assert np.array_equal(r_a[k], x[k]), k
    off = a["call_off"].astype(np.int64)
    assert (np.diff(off) >= 0).all() and off[0] == 0 and off[-1] == len(a["calls"]) == a["n_calls"]
    v = a["calls"]["var"]
    starts = off[:-1][np.diff(off) > 0]
    inner = np.ones(len(v), bool)
    inner[starts] = False
    assert (np.diff(v.astype(np.int64))[inner[1:]] > 0).all(), "calls of a read must be in ascending variant order"
    assert set(np.unique(a["calls"]["allele"])) <= {0, 1}
    assert (r_a["ps"] != 0).sum() > 0.9 * c.n_var                # 30x: nearly everything phases
    ctx.close()
