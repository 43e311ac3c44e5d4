"""`somatic_haplotag` of the C++ host (longphase-s_b200/host/somatic_host.cpp) against the UNMODIFIED reference binary on the same
tumor / normal files: tagged tumor BAM (HP:Z / PS:i / PQ:i, uncompressed byte stream) and <prefix>_purity.out must be identical.

CPU: the host's own stages (options, both VCF loaders and their union map, BAM packing for the three passes, purity report,
calling stage, BAM writer) run for real; the ORACLE stands in for the device in the three batch calls (lps_extract_normal,
lps_extract_tumor, lps_somatic_tag_reads).  -m gpu: the real binary, device and all."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from . import host_cli as hc
from .test_host_cli import needs_host, needs_ref, run_in

po = pytest.importorskip("oracle.pyoracle")
ffi = importlib.import_module("longphase_s_b200._ffi")

_data = {}


def dataset(tmp_path_factory):
    if "files" not in _data:
        d = str(tmp_path_factory.mktemp("hostcli_somatic"))
        pairs = []
        for name, kw, dn, dt, purity in (("chrA", dict(seed=421, contig_len=400_000, indel_frac=0.15, somatic_rate=1 / 3000.0), 25, 50, 0.6),
                                        ("chrB", dict(seed=422, contig_len=250_000, indel_frac=0.1, somatic_rate=1 / 2500.0, supp_frac=0.15), 25, 45, 0.6)):
            cn = hc.synth.Contig(**kw, depth=dn, purity=0.0, read_seed=1000 + kw["seed"])
            ct = hc.synth.Contig(**kw, depth=dt, purity=purity, read_seed=2000 + kw["seed"])
            pairs.append((name, cn, ct))
        files = hc.write_somatic_dataset(d, pairs)
        # the phased NORMAL VCF both programs read: the reference's own `phase` on the normal BAM
        run_in(os.path.join(d, "phase"), [hc.REF_BIN, "phase", "-s", files["germline_vcf"], "-b", files["normal_bam"], "-r", files["fasta"], "-o", "normal",
                                         "--ont", "--indels", "-t", "2"])
        files["normal_vcf"] = os.path.join(d, "phase", "normal.vcf")
        files["dir"] = d
        _data["files"] = files
    return _data["files"]


def som_args(files, extra):
    return ["somatic_haplotag", "-s", files["normal_vcf"], "-b", files["normal_bam"], "--tumor-snv-file", files["tumor_vcf"], "--tumor-bam-file",
            files["tumor_bam"], "-r", files["fasta"], "-o", "som", "-t", "2"] + extra


def som_lib():
    lib = hc.host_lib()
    vp = C.c_void_p
    lib.lpsh_som_open.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp)]
    lib.lpsh_som_n_contigs.argtypes = [vp]
    lib.lpsh_som_params.argtypes = [vp, C.c_int, C.POINTER(ffi.LpsTagParams)]
    lib.lpsh_som_pack.argtypes = [vp, C.c_int, C.c_int, C.POINTER(hc.LpshPacked), C.POINTER(ffi.LpsTumorVariants)]
    lib.lpsh_som_set_extract.argtypes = [vp, C.c_int, C.c_int, C.POINTER(ffi.LpsExtractResult)]
    lib.lpsh_som_call.argtypes = [vp]
    lib.lpsh_som_estimate.argtypes = [vp]
    lib.lpsh_som_purity.argtypes = [vp]
    lib.lpsh_som_purity.restype = C.c_double
    lib.lpsh_som_n_somatic.argtypes = [vp]
    lib.lpsh_som_n_somatic.restype = C.c_int64
    lib.lpsh_som_tag_begin.argtypes = [vp]
    lib.lpsh_som_tag_pack.argtypes = [vp, C.c_int, C.POINTER(hc.LpshPacked), C.POINTER(ffi.LpsTumorVariants)]
    lib.lpsh_som_tag_emit.argtypes = [vp, C.c_int, C.POINTER(ffi.LpsSomaticTagResult)]
    lib.lpsh_som_tag_end.argtypes = [vp]
    lib.lpsh_som_close.argtypes = [vp]
    lib.lpsh_som_tag_run_with.argtypes = [vp, SOM_JUDGE_FN, vp]
    return lib


SOM_JUDGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(hc.LpshPacked), C.POINTER(ffi.LpsTumorVariants), C.POINTER(ffi.LpsSomaticTagResult))


def union_contig(pk, tv):
    """packed contig + lps_tumor_variants -> the union-contig object the somatic oracle takes (numpy copies)."""
    g = ffi.as_np
    c = hc.packed_contig(pk)
    n = tv.n
    c.nor_present, c.tum_present = g(tv.nor_present, n, np.uint8), g(tv.tum_present, n, np.uint8)
    c.tum_ref0, c.tum_alt0 = g(tv.ref0, n, np.uint8), g(tv.alt0, n, np.uint8)
    c.tum_ref_len, c.tum_alt_len = g(tv.ref_len, n, np.uint16), g(tv.alt_len, n, np.uint16)
    c.tum_gt, c.tum_hp1_is_alt, c.tum_ps = g(tv.gt_kind, n, np.uint8), g(tv.hp1_is_alt, n, np.uint8), g(tv.ps, n, np.int32)
    c.is_somatic, c.derive_hp = g(tv.is_somatic, n, np.uint8), g(tv.derive_hp, n, np.int8)
    return c


def extract_struct(o, keep):
    """po.OracleSomatic -> lps_extract_result (the arrays stay referenced through `keep`)."""
    P = ffi.ptr

    def arr(x, dt):
        a = np.ascontiguousarray(x, dt)
        keep.append(a)
        return a
    reads = ffi.LpsReadTags(n_reads=o.n_reads, category=P(arr(o.category, np.uint8), ffi.u8p), read_hp=P(arr(o.read_hp, np.int8), ffi.i8p),
                            ps=P(arr(o.ps, np.int32), ffi.i32p), pq=P(arr(o.pq, np.int32), ffi.i32p), h1=P(arr(o.h1, np.int32), ffi.i32p),
                            h2=P(arr(o.h2, np.int32), ffi.i32p), h3=P(arr(o.h3, np.int32), ffi.i32p), n_ps=P(arr(o.n_ps, np.uint8), ffi.u8p),
                            end_pos=P(arr(o.end_pos, np.int32), ffi.i32p), read_len=P(arr(o.read_len, np.int32), ffi.i32p))
    calls = arr(o.calls, ffi.CALL_DTYPE)
    return ffi.LpsExtractResult(n_tum=o.n_tum, tum_var=P(arr(o.tum_var, np.int32), ffi.i32p), pos_base=P(arr(o.pos_base, np.int32), ffi.i32p),
                                read_hp_count=P(arr(o.read_hp_count, np.int32), ffi.i32p), reads=reads,
                                somatic_read_hp_count=P(arr(o.somatic_read_hp_count, np.int32), ffi.i32p),
                                case_count=P(arr(o.case_count, np.int32), ffi.i32p), allele_count=P(arr(o.allele_count, np.int32), ffi.i32p),
                                window_hist=P(arr(o.window_hist, np.int32), ffi.i32p), n_window_items=o.n_window_items,
                                ratios_f=P(arr(o.ratios_f, np.float32), ffi.f32p), ratios_d=P(arr(o.ratios_d, np.float64), ffi.f64p),
                                case_read_count=P(arr(o.case_read_count, np.int32), ffi.i32p), n_calls=len(calls),
                                call_off=P(arr(o.call_off, np.uint64), ffi.u64p), calls=calls.ctypes.data_as(C.POINTER(ffi.LpsCall)))


def tag_struct(o, c, tp, keep):
    """po.OracleSomatic(mode somatic_tag) -> lps_somatic_tag_result, ReadStatistics reduced as lps_somatic_tag_reads does."""
    P = ffi.ptr

    def arr(x, dt):
        a = np.ascontiguousarray(x, dt)
        keep.append(a)
        return a
    r = ffi.LpsSomaticTagResult()
    r.reads = ffi.LpsReadTags(n_reads=o.n_reads, category=P(arr(o.category, np.uint8), ffi.u8p), read_hp=P(arr(o.read_hp, np.int8), ffi.i8p),
                              ps=P(arr(o.ps, np.int32), ffi.i32p), pq=P(arr(o.pq, np.int32), ffi.i32p))
    r.total_alignment = o.n_reads
    return r


def oracle_somatic_through_host(files, extra, cwd, chunk=300, pipelined=False):
    lib = som_lib()
    os.makedirs(cwd, exist_ok=True)
    old = os.getcwd()
    os.chdir(cwd)
    os.environ["LPS_TAG_CHUNK"] = str(chunk)
    try:
        h = C.c_void_p()
        n, av = hc.argv(som_args(files, extra))
        assert lib.lpsh_som_open(n, av, C.byref(h)) == 0, lib.lpsh_last_error()
        xp, tp = ffi.LpsTagParams(), ffi.LpsTagParams()
        lib.lpsh_som_params(h, 0, C.byref(xp))
        lib.lpsh_som_params(h, 1, C.byref(tp))
        assert xp.mapq_filter == 0 and tp.mapq_filter == 1
        nc = lib.lpsh_som_n_contigs(h)
        for which, mode in ((0, "extract_normal"), (1, "extract_tumor")):
            for i in range(nc):
                pk, tv = hc.LpshPacked(), ffi.LpsTumorVariants()
                assert lib.lpsh_som_pack(h, i, which, C.byref(pk), C.byref(tv)) == 0, lib.lpsh_last_error()
                c = union_contig(pk, tv)
                o = po.OracleSomatic(c, xp, mode)
                assert o.rc == 0
                keep = []
                r = extract_struct(o, keep)
                assert lib.lpsh_som_set_extract(h, i, which, C.byref(r)) == 0
        assert lib.lpsh_som_call(h) == 0, lib.lpsh_last_error()
        info = dict(purity=lib.lpsh_som_purity(h), n_somatic=lib.lpsh_som_n_somatic(h), chunks=0, h3_reads=0)
        if pipelined:     # the host's own reader-thread / writer pipeline with the oracle as the judge of every chunk
            state = dict(keep=None)

            def judge(user, contig, pk, tvp, out):
                try:
                    c = union_contig(pk.contents, tvp.contents)
                    o = po.OracleSomatic(c, tp, "somatic_tag")
                    info["h3_reads"] += int((o.read_hp >= 3).sum())
                    info["chunks"] += 1
                    keep = []
                    r = tag_struct(o, c, tp, keep)
                    C.memmove(out, C.byref(r), C.sizeof(r))
                    state["keep"] = keep
                    return 0 if o.rc == 0 else -1
                except Exception as e:  # noqa: BLE001
                    print("judge failed:", e)
                    return -1
            cb = SOM_JUDGE_FN(judge)
            assert lib.lpsh_som_tag_run_with(h, cb, None) == 0, lib.lpsh_last_error()
        else:
            assert lib.lpsh_som_tag_begin(h) == 0, lib.lpsh_last_error()
            for i in range(nc):
                while True:
                    pk, tv = hc.LpshPacked(), ffi.LpsTumorVariants()
                    got = lib.lpsh_som_tag_pack(h, i, C.byref(pk), C.byref(tv))
                    assert got >= 0, lib.lpsh_last_error()
                    if got == 0:
                        break
                    info["chunks"] += 1
                    c = union_contig(pk, tv)
                    o = po.OracleSomatic(c, tp, "somatic_tag")
                    assert o.rc == 0
                    info["h3_reads"] += int((o.read_hp >= 3).sum())
                    keep = []
                    r = tag_struct(o, c, tp, keep)
                    assert lib.lpsh_som_tag_emit(h, i, C.byref(r)) == 0, lib.lpsh_last_error()
            assert lib.lpsh_som_tag_end(h) == 0
        lib.lpsh_som_close(h)
        return info
    finally:
        os.environ.pop("LPS_TAG_CHUNK", None)
        os.chdir(old)


SOM_VARIANTS = [["--output-somatic-vcf", "--somatic-calling-log"], ["--tumor-purity", "0.35", "--tagSupplementary", "-q", "20", "--somatic-calling-log"],
                ["--disableFilter", "-p", "0.7"], ["--region", "chrA:40000-260000", "--tumor-purity", "0.6", "--somatic-calling-log"]]


@needs_host
@needs_ref
@pytest.mark.parametrize("extra", SOM_VARIANTS)
def test_somatic_host_files_match_reference(tmp_path_factory, tmp_path, extra):
    files = dataset(tmp_path_factory)
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + som_args(files, extra))
    info = oracle_somatic_through_host(files, extra, str(tmp_path / "own"), pipelined=(extra != SOM_VARIANTS[1]))
    assert info["chunks"] >= (2 if "--region" in extra else 4) and info["n_somatic"] > 20 and info["h3_reads"] > 50, info
    ref, own = hc.bam_payload(str(tmp_path / "ref" / "som.bam")), hc.bam_payload(str(tmp_path / "own" / "som.bam"))
    assert b"HPZ" in own, "no HP:Z tag was written"
    assert own == ref, "tagged tumor BAM differs from the reference's (uncompressed byte stream)"
    if "--output-somatic-vcf" in extra:
        sc = open(tmp_path / "own" / "som_sc.vcf").read()
        assert hc.strip_commandline(sc) == hc.strip_commandline(open(tmp_path / "ref" / "som_sc.vcf").read())
        assert "\tLowQual\t" in sc and "\tPASS\t" in sc and "##longphase_s_version=" in sc
    if "--somatic-calling-log" in extra:
        # BASELINE.md's gate file: one row of 65 columns per position called somatic, every number as the reference prints it
        log = open(tmp_path / "own" / "som_somatic_var.out").read()
        assert log == open(tmp_path / "ref" / "som_somatic_var.out").read()
        rows = [l for l in log.splitlines() if l and not l.startswith("#")]
        assert len(rows) == info["n_somatic"] and all(len(r.split("\t")) >= 65 for r in rows)
    if "--tumor-purity" not in extra:
        assert 0.0 < info["purity"] <= 1.0
        assert open(tmp_path / "own" / "som_purity.out").read() == open(tmp_path / "ref" / "som_purity.out").read()
    else:
        assert not os.path.exists(tmp_path / "own" / "som_purity.out") and not os.path.exists(tmp_path / "ref" / "som_purity.out")


@needs_host
@needs_ref
def test_estimate_purity_host_matches_reference(tmp_path_factory, tmp_path):
    """`estimate_purity` (PurityEstimation.cpp): the two extract passes with its own defaults (-q 20, supplementary included) and the
    purity stage; <prefix>_purity.out identical to the reference's."""
    files = dataset(tmp_path_factory)
    args = ["estimate_purity", "-s", files["normal_vcf"], "-b", files["normal_bam"], "--tumor-snv-file", files["tumor_vcf"], "--tumor-bam-file",
            files["tumor_bam"], "-r", files["fasta"], "-o", "pur", "-t", "2"]
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + args)
    lib = som_lib()
    os.makedirs(tmp_path / "own")
    old = os.getcwd()
    os.chdir(str(tmp_path / "own"))
    try:
        h = C.c_void_p()
        n, av = hc.argv(args)
        assert lib.lpsh_som_open(n, av, C.byref(h)) == 0, lib.lpsh_last_error()
        xp = ffi.LpsTagParams()
        lib.lpsh_som_params(h, 0, C.byref(xp))
        assert xp.mapping_quality == 20 and xp.tag_supplementary == 1 and xp.mapq_filter == 0
        for which, mode in ((0, "extract_normal"), (1, "extract_tumor")):
            for i in range(lib.lpsh_som_n_contigs(h)):
                pk, tv = hc.LpshPacked(), ffi.LpsTumorVariants()
                assert lib.lpsh_som_pack(h, i, which, C.byref(pk), C.byref(tv)) == 0, lib.lpsh_last_error()
                o = po.OracleSomatic(union_contig(pk, tv), xp, mode)
                keep = []
                r = extract_struct(o, keep)
                assert lib.lpsh_som_set_extract(h, i, which, C.byref(r)) == 0
        assert lib.lpsh_som_estimate(h) == 0, lib.lpsh_last_error()
        purity = lib.lpsh_som_purity(h)
        lib.lpsh_som_close(h)
    finally:
        os.chdir(old)
    assert 0.0 < purity <= 1.0
    own = open(tmp_path / "own" / "pur_purity.out").read()
    assert own == open(tmp_path / "ref" / "pur_purity.out").read()
    assert ("Tumor purity: %g" % purity) in own or "Tumor purity: " in own


@needs_host
@needs_ref
def test_somatic_over_the_batched_inflate_reader(tmp_path_factory, tmp_path):
    """LPS_GPU_INFLATE=1 with zlib as the inflater: the extract passes pack straight from the inflated stream, the tagging pass builds its
    bam1_t records from it; same tagged tumor BAM and purity report as the reference."""
    from .test_host_cli import zlib_inflater
    files = dataset(tmp_path_factory)
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + som_args(files, []))
    with zlib_inflater() as stats:
        info = oracle_somatic_through_host(files, [], str(tmp_path / "own"), pipelined=True)
    assert stats["calls"] == 6 and info["h3_reads"] > 50          # 2 contigs x (normal pass, tumor pass, tagging pass)
    assert hc.bam_payload(str(tmp_path / "own" / "som.bam")) == hc.bam_payload(str(tmp_path / "ref" / "som.bam"))
    assert open(tmp_path / "own" / "som_purity.out").read() == open(tmp_path / "ref" / "som_purity.out").read()


@needs_host
@needs_ref
@pytest.mark.parametrize("extra", [[], ["--region", "chrA:40000-260000", "--tumor-purity", "0.6"]])
def test_somatic_extract_passes_with_split_readers(tmp_path_factory, tmp_path, extra):
    """LPS_READ_SPLIT=5: the whole-contig batches of the two extract passes read by five readers on slices of the region (also a user region
    that starts inside the contig: the first slice takes the alignments that start before it); same files as the reference."""
    files = dataset(tmp_path_factory)
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + som_args(files, extra))
    os.environ["LPS_READ_SPLIT"] = "5"
    try:
        oracle_somatic_through_host(files, extra, str(tmp_path / "own"), pipelined=True)
    finally:
        os.environ.pop("LPS_READ_SPLIT", None)
    assert hc.bam_payload(str(tmp_path / "own" / "som.bam")) == hc.bam_payload(str(tmp_path / "ref" / "som.bam"))
    if not extra:
        assert open(tmp_path / "own" / "som_purity.out").read() == open(tmp_path / "ref" / "som_purity.out").read()


@needs_host
def test_somatic_binaries_need_a_gpu(tmp_path_factory, tmp_path):
    """No CPU fallback: `somatic_haplotag` and `estimate_purity` stop with an error when no CUDA device is usable."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    files = dataset(tmp_path_factory)
    for sub in ("somatic_haplotag", "estimate_purity"):
        args = [sub] + som_args(files, [])[1:]
        p = subprocess.run([hc.HOST_BIN] + args, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 1 and "no usable CUDA device" in p.stderr, sub
        assert not os.path.exists(tmp_path / "som.bam") and not os.path.exists(tmp_path / "som_purity.out")


@needs_host
def test_somatic_host_rejects_bad_options(capfd):
    lib = som_lib()
    h = C.c_void_p()
    n, av = hc.argv(["somatic_haplotag", "-s", "/nonexistent.vcf", "--tumor-purity", "1.5", "--log"])   # --log: per-read table, not rebuilt
    assert lib.lpsh_som_open(n, av, C.byref(h)) == 1 and not h.value
    err = capfd.readouterr().err
    assert "SNP file" in err and "invalid tumor purity" in err and "not available in this build" in err


@pytest.mark.gpu
@needs_host
@needs_ref
def test_gpu_cli_somatic_haplotag_matches_reference(tmp_path_factory, tmp_path):
    files = dataset(tmp_path_factory)
    for k, extra in enumerate([["--somatic-calling-log"], ["--tumor-purity", "0.35", "--tagSupplementary", "-q", "20", "--somatic-calling-log"]]):
        run_in(str(tmp_path / f"ref{k}"), [hc.REF_BIN] + som_args(files, extra))
        os.makedirs(tmp_path / f"own{k}", exist_ok=True)
        p = subprocess.run([hc.HOST_BIN] + som_args(files, extra), cwd=str(tmp_path / f"own{k}"), env=dict(os.environ, LPS_TAG_CHUNK="2500"),
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        assert hc.bam_payload(str(tmp_path / f"own{k}" / "som.bam")) == hc.bam_payload(str(tmp_path / f"ref{k}" / "som.bam")), extra
        assert open(tmp_path / f"own{k}" / "som_somatic_var.out").read() == open(tmp_path / f"ref{k}" / "som_somatic_var.out").read(), extra
        if "--tumor-purity" not in extra:
            assert open(tmp_path / f"own{k}" / "som_purity.out").read() == open(tmp_path / f"ref{k}" / "som_purity.out").read()
