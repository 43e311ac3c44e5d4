"""The C++ host above the C ABI (longphase-s_b200/host): `phase` and `haplotag` with the reference's command line and files.

CPU tests drive every host stage that needs no GPU — options, VCF / FASTA / BAM loading and SoA packing with htslib, the
phased-VCF writer, the tagged-BAM writer and the --log table — with the ORACLE standing in for the device between pack and
write, and compare the written files with those of the UNMODIFIED reference binary (oracle/_ref/longphase-s) run on the same
inputs: the phased VCF must be identical except for its ##commandline line, the tagged BAM's uncompressed byte stream and
the .out table must be identical.  The -m gpu tests run the real binary (`longphase-s-b200`), device and all, the same way."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from . import host_cli as hc

po = pytest.importorskip("oracle.pyoracle")
ffi = importlib.import_module("longphase_s_b200._ffi")

needs_host = pytest.mark.skipif(not os.path.exists(hc.HOST_LIB), reason="liblps_host.so not built (needs the reference's htslib sources)")
needs_ref = pytest.mark.skipif(not (os.path.exists(hc.REF_BIN) and os.path.exists(hc.MKBAM)), reason="reference binary not built")

_data = {}


def dataset(tmp_path_factory, kind):
    if kind not in _data:
        d = str(tmp_path_factory.mktemp("hostcli_" + kind))
        a = hc.synth.Contig(seed=201, contig_len=260_000, indel_frac=0.12, depth=18.0, mean_len=9_000.0)
        b = hc.synth.Contig(seed=202, contig_len=180_000, indel_frac=0.05, depth=14.0, mean_len=7_000.0, supp_frac=0.2)
        e = hc.synth.Contig(seed=203, contig_len=30_000, depth=4.0, mean_len=3_000.0)
        files = hc.write_dataset(d, [("chrA", a, True), ("chrEmpty", e, False), ("chrB", b, True)],
                                 with_ps=(kind == "ps_gz"), gz=(kind == "ps_gz"))
        files["dir"] = d
        _data[kind] = files
    return _data[kind]


def run_in(cwd, cmd):
    os.makedirs(cwd, exist_ok=True)
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, (cmd, p.stdout[-2000:], p.stderr[-2000:])
    return p


def phase_args(files, extra):
    return ["phase", "-s", files["vcf"], "-b", files["bam"], "-r", files["fasta"], "-o", "out", "-t", "2"] + extra


def oracle_phase_through_host(files, extra, cwd):
    """lpsh_phase_open -> per contig: lpsh_phase_pack -> ORACLE -> lpsh_phase_set_result -> lpsh_phase_write_result."""
    lib = hc.host_lib()
    os.makedirs(cwd, exist_ok=True)
    old = os.getcwd()
    os.chdir(cwd)
    try:
        h = C.c_void_p()
        n, av = hc.argv(phase_args(files, extra))
        assert lib.lpsh_phase_open(n, av, C.byref(h)) == 0, lib.lpsh_last_error()
        p = ffi.LpsPhaseParams()
        assert lib.lpsh_phase_params(h, C.byref(p)) == 0
        seen = {}
        for i in range(lib.lpsh_phase_n_contigs(h)):
            name = lib.lpsh_phase_contig_name(h, i).decode()
            if lib.lpsh_phase_last_variant(h, i) == -1:
                continue
            pk = hc.LpshPacked()
            assert lib.lpsh_phase_pack(h, i, C.byref(pk)) == 0, lib.lpsh_last_error()
            c = hc.packed_contig(pk)
            lib.lpsh_phase_release(h, i)
            seen[name] = c
            if c.n_reads == 0:
                continue
            orc = po.OraclePhase(c, p)
            assert orc.rc == 0
            ps, hap = np.ascontiguousarray(orc.ps, np.int32), np.ascontiguousarray(orc.hap_ref, np.int8)
            assert lib.lpsh_phase_set_result(h, i, c.n_var, ffi.ptr(ps, ffi.i32p), ffi.ptr(hap, ffi.i8p)) == 0
        assert lib.lpsh_phase_write_result(h) == 0
        lib.lpsh_phase_close(h)
        return seen
    finally:
        os.chdir(old)


PHASE_VARIANTS = [("plain", ["--ont"]), ("plain", ["--ont", "--indels"]), ("plain", ["--pb", "--indels", "--indelQuality", "30", "-q", "20", "-a", "20"]),
                  ("ps_gz", ["--ont", "--indels"])]


@needs_host
@needs_ref
@pytest.mark.parametrize("kind,extra", PHASE_VARIANTS)
def test_phase_host_files_match_reference(tmp_path_factory, tmp_path, kind, extra):
    files = dataset(tmp_path_factory, kind)
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + phase_args(files, extra))
    seen = oracle_phase_through_host(files, extra, str(tmp_path / "own"))
    assert set(seen) == {"chrA", "chrB"} and all(c.n_reads > 100 for c in seen.values())
    ref = open(tmp_path / "ref" / "out.vcf").read()
    own = open(tmp_path / "own" / "out.vcf").read()
    assert "|" in own and ":PS" in own
    assert hc.strip_commandline(own) == hc.strip_commandline(ref)
    n_phased = sum(1 for ln in own.split("\n") if ln and ln[0] != "#" and "|" in ln.split("\t")[9])
    assert n_phased > 100
    if "--indelQuality" in extra:
        assert open(tmp_path / "own" / "out_removed_indels.log").read() == open(tmp_path / "ref" / "out_removed_indels.log").read()
        assert "INDEL_QUAL_FILTERED" in own


@needs_host
@needs_ref
def test_phase_host_with_two_bam_files(tmp_path_factory, tmp_path):
    """Repeated -b: the reference appends the alignments of every file to one read set per contig (ParsingBam.cpp:1251-1299); the same
    BAM given twice doubles every read name, so the merge-by-name and the overlap filter of addEdge see pairs everywhere."""
    files = dataset(tmp_path_factory, "plain")
    two = dict(files)
    args = lambda f, extra: ["phase", "-s", f["vcf"], "-b", f["bam"], "-b", f["bam"], "-r", f["fasta"], "-o", "out", "-t", "2"] + extra   # noqa: E731
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + args(two, ["--ont"]))
    import tests.test_host_cli as me
    saved = me.phase_args
    me.phase_args = args
    try:
        seen = oracle_phase_through_host(two, ["--ont"], str(tmp_path / "own"))
    finally:
        me.phase_args = saved
    names = seen["chrA"].read_names
    assert seen["chrA"].n_reads > 800 and all(names.count(x) % 2 == 0 for x in set(names[:50]))      # every alignment came in twice
    assert hc.strip_commandline(open(tmp_path / "own" / "out.vcf").read()) == hc.strip_commandline(open(tmp_path / "ref" / "out.vcf").read())


INFLATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(ffi.LpsBgzfBlock), C.c_uint64, C.POINTER(C.c_uint8), C.c_uint64)


class zlib_inflater:
    """with zlib_inflater() as stats: LPS_GPU_INFLATE=1 with zlib installed as the batched inflater (lpsh_set_inflater), so that the
    readers around the device call run on the CPU."""

    def __enter__(self):
        import zlib
        self.stats = stats = dict(calls=0, blocks=0)

        def inflate(user, data, n_bytes, blocks, n_blocks, out, out_cap):
            try:
                src = C.string_at(data, n_bytes)
                for k in range(n_blocks):
                    b = blocks[k]
                    raw = zlib.decompress(src[b.comp_off:b.comp_off + b.comp_len], -15)
                    assert len(raw) == b.out_len and b.out_off + b.out_len <= out_cap and zlib.crc32(raw) == b.crc32
                    C.memmove(C.addressof(out.contents) + b.out_off, raw, len(raw))
                stats["calls"] += 1
                stats["blocks"] += n_blocks
                return 0
            except Exception as e:  # noqa: BLE001
                print("inflate hook failed:", e)
                return -1
        self.cb = INFLATE_FN(inflate)
        lib = hc.host_lib()
        lib.lpsh_set_inflater.argtypes = [INFLATE_FN, C.c_void_p]
        lib.lpsh_set_inflater(self.cb, None)
        os.environ["LPS_GPU_INFLATE"] = "1"
        return stats

    def __exit__(self, *exc):
        os.environ.pop("LPS_GPU_INFLATE", None)
        hc.host_lib().lpsh_set_inflater(C.cast(None, INFLATE_FN), None)
        return False


@needs_host
def test_batched_inflate_reader_packs_what_htslib_packs(tmp_path_factory, tmp_path):
    """LPS_GPU_INFLATE=1: the region's compressed bytes -> lps_bgzf_scan -> one batched inflate -> records parsed from memory.  With zlib
    installed as the inflater (lpsh_set_inflater) everything around the device call runs here: index chunks to file range, member
    table, record walk with hts_itr_next's acceptance test, raw-record packing.  Every array must equal the htslib reader's."""
    files = dataset(tmp_path_factory, "plain")
    plain = oracle_phase_through_host(files, ["--ont", "--indels"], str(tmp_path / "a"))
    with zlib_inflater() as stats:
        batched = oracle_phase_through_host(files, ["--ont", "--indels"], str(tmp_path / "b"))
    assert stats["calls"] == 2 and stats["blocks"] > 100
    for name in ("chrA", "chrB"):
        a, b = plain[name], batched[name]
        assert a.n_reads == b.n_reads and a.read_names == b.read_names
        for k in ("ref_start", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "flag", "mapq", "name_rank", "cigar", "seq4", "qual"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (name, k)
    assert open(tmp_path / "a" / "out.vcf").read() == open(tmp_path / "b" / "out.vcf").read()


@needs_host
@needs_ref
@pytest.mark.parametrize("extra", [["--log"], ["--log", "--region", "chrB:20000-120000"]])
def test_haplotag_over_the_batched_inflate_reader(tmp_path_factory, tmp_path, extra):
    """The tagging pass reading its records from the batched-inflate stream (InflatedRegion::to_bam1 = what bam_read1 leaves in a bam1_t,
    bin recomputed, name padding): tagged BAM and .out identical to the reference's, old HP / PS tags of the input replaced, other aux kept."""
    files = dataset(tmp_path_factory, "plain")
    if "phased_vcf" not in files:
        d = os.path.join(files["dir"], "phase_ref")
        run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
        files["phased_vcf"] = os.path.join(d, "out.vcf")
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, files["phased_vcf"], extra))
    with zlib_inflater() as stats:
        oracle_tag_pipelined(files, files["phased_vcf"], extra, str(tmp_path / "own"), 150)
    assert stats["calls"] == (1 if "--region" in extra else 3)
    own = hc.bam_payload(str(tmp_path / "own" / "tagged.bam"))
    assert own == hc.bam_payload(str(tmp_path / "ref" / "tagged.bam")) and b"XXZkeep" in own
    assert open(tmp_path / "own" / "tagged.out").read() == open(tmp_path / "ref" / "tagged.out").read()


@needs_host
@needs_ref
def test_phase_vcf_rewrite_quirks(tmp_path):
    """The text side of `phase` alone: 400 records on a contig without reads, with every FORMAT order, an old PS at any place, a key that
    merely contains "PS", phased / unphased / missing / multi-allelic genotypes, indels, QUAL and FILTER variants.  Loader and writer
    must rewrite each line exactly as SnpParser::writeLine does (colon counting, PS removal, 1|0 -> 0/1, :PS append, quality filter)."""
    import random
    d = str(tmp_path)
    a = hc.synth.Contig(seed=301, contig_len=30_000, depth=5.0, mean_len=3_000.0)
    e = hc.synth.Contig(seed=302, contig_len=20_000, depth=0.2, mean_len=2_000.0)
    e.n_reads = 0                                            # no alignment on chrF: its records are only rewritten
    files = hc.write_dataset(d, [("chrA", a, True), ("chrF", e, False)], fast_bam=True)
    rng = random.Random(7)
    ref = e.ref.decode()
    gts = ["0/1", "1/0", "0|1", "1|0", "1/1", "0/0", "./.", "1|1", "0|0", "1/2", "0/2", "1|2", "2|1"]
    fmts = [("GT", "{gt}"), ("GT:PS", "{gt}:{ps}"), ("GT:DP:PS", "{gt}:30:{ps}"), ("PS:GT", "{ps}:{gt}"), ("DP:GT:PS:GQ", "30:{gt}:{ps}:50"),
            ("GT:GQ:DP", "{gt}:40:20"), ("DP:GT", "12:{gt}"), ("GT:PSX", "{gt}:9"), ("GT:AD:PS", "{gt}:3,4:{ps}")]
    lines, pos = [], 100
    for _ in range(400):
        pos += rng.randint(3, 40)
        r = ref[pos - 1].upper()
        kind = rng.random()
        if kind < 0.6:
            R, alt = r, rng.choice([x for x in "ACGT" if x != r])
        elif kind < 0.75:
            R, alt = r, r + "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 4)))
        elif kind < 0.9:
            R, alt = ref[pos - 1:pos + rng.randint(1, 4)].upper(), r
        else:
            R, alt = r, rng.choice([x for x in "ACGT" if x != r]) + "," + rng.choice("ACGT")
        fmt, smp = rng.choice(fmts)
        lines.append("chrF\t%d\t.\t%s\t%s\t%s\t%s\tDP=3\t%s\t%s" % (pos, R, alt, rng.choice(["50", ".", "7.5", "30"]), rng.choice(["PASS", ".", "q10"]), fmt,
                                                                        smp.format(gt=rng.choice(gts), ps=rng.randint(1, 5) * 100)))
    text = open(files["vcf"]).read().split("\n")
    hdr, body = [ln for ln in text if ln.startswith("#")], [ln for ln in text if ln and not ln.startswith("#")]
    for extra_hdr in ('##FORMAT=<ID=PS,Number=1,Type=Integer,Description="ps">', '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="ad">',
                      '##FORMAT=<ID=PSX,Number=1,Type=Integer,Description="x">', '##FILTER=<ID=q10,Description="q">'):
        hdr.insert(-1, extra_hdr)
    open(files["vcf"], "w").write("\n".join(hdr + body + lines) + "\n")
    for k, extra in enumerate((["--ont"], ["--ont", "--indels"], ["--pb", "--indels", "--indelQuality", "20"])):
        run_in(os.path.join(d, "ref%d" % k), [hc.REF_BIN] + phase_args(files, extra))
        oracle_phase_through_host(files, extra, os.path.join(d, "own%d" % k))
        ref_text = hc.strip_commandline(open(os.path.join(d, "ref%d" % k, "out.vcf")).read())
        assert hc.strip_commandline(open(os.path.join(d, "own%d" % k, "out.vcf")).read()) == ref_text, extra
        assert ref_text.count("chrF\t") == 400


@needs_host
@needs_ref
def test_phase_deepsomatic_preprocessing(tmp_path):
    """--deepsomatic_output: GERMLINE records only, genotype re-derived from AD (or VAF) before phasing (ParsingBam.cpp:651-834);
    <prefix>_preprocessed.vcf and the phased VCF equal the reference's."""
    import random
    d = str(tmp_path)
    a = hc.synth.Contig(seed=321, contig_len=150_000, indel_frac=0.1, depth=14.0, mean_len=7_000.0)
    files = hc.write_dataset(d, [("chrA", a, True)], fast_bam=True)
    rng, out = random.Random(11), []
    for ln in open(files["vcf"]).read().split("\n"):
        if ln.startswith("#CHROM"):
            out += ['##FILTER=<ID=GERMLINE,Description="g">', '##FILTER=<ID=SOMATIC,Description="s">', '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="ad">',
                    '##FORMAT=<ID=VAF,Number=A,Type=Float,Description="vaf">']
        if ln and not ln.startswith("#"):
            t = ln.split("\t")
            t[6] = rng.choice(["GERMLINE", "GERMLINE", "GERMLINE", "SOMATIC", "PASS", "GERMLINE;LowQ"])
            u, ref_n, alt_n = rng.random(), rng.randint(0, 30), rng.randint(0, 30)
            vaf = "%.3f" % rng.random()
            if u < 0.15 and len(t[3]) == 1 and len(t[4]) == 1:          # second ALT allele, three AD values
                t[4] += "," + rng.choice([x for x in "ACGT" if x != t[4]])
                t[8], t[9] = "GT:AD:VAF", "0/1:%d,%d,%d:%s,%s" % (ref_n, alt_n, rng.randint(0, 9), vaf, "0.05")
            elif u < 0.5:
                t[8], t[9] = "GT:AD:VAF", "0/1:%d,%d:%s" % (ref_n, alt_n, vaf)
            elif u < 0.65:
                t[8], t[9] = "GT:VAF", "1/1:%s" % vaf
            elif u < 0.75:
                t[8], t[9] = "GT:AD:VAF", "0/1:.,.:%s" % vaf                # AD unusable -> VAF
            elif u < 0.85:
                t[8], t[9] = "GT:AD", "0/0:%d" % ref_n                    # wrong AD arity and no VAF: genotype kept
            else:
                t[8], t[9] = "GT:DP:VAF", "0|1:30:."
            ln = "\t".join(t)
        out.append(ln)
    open(files["vcf"], "w").write("\n".join(out))
    extra = ["--ont", "--indels", "--deepsomatic_output"]
    run_in(os.path.join(d, "ref"), [hc.REF_BIN] + phase_args(files, extra))
    oracle_phase_through_host(files, extra, os.path.join(d, "own"))
    pre = open(os.path.join(d, "own", "out_preprocessed.vcf")).read()
    assert pre == open(os.path.join(d, "ref", "out_preprocessed.vcf")).read()
    assert "SOMATIC" not in pre.split("#CHROM")[1] and "\t1/1:" in pre and "\t0/0:" in pre and "\t0/2:" in pre + "\t0/2:"
    own = hc.strip_commandline(open(os.path.join(d, "own", "out.vcf")).read())
    assert own == hc.strip_commandline(open(os.path.join(d, "ref", "out.vcf")).read()) and own.count("|") > 20


@needs_host
@pytest.mark.parametrize("readers", [3, 7])
def test_split_region_readers_pack_the_same_batch(tmp_path_factory, tmp_path, readers):
    """LPS_READ_SPLIT=K (what `phase -t N` does by itself when it has fewer contigs than threads): K readers on slices of the contig's
    region, a record belonging to the slice its start lies in; the packed batch must be the single iterator's, array for array."""
    files = dataset(tmp_path_factory, "plain")
    plain = oracle_phase_through_host(files, ["--ont", "--indels"], str(tmp_path / "a"))
    os.environ["LPS_READ_SPLIT"] = str(readers)
    try:
        split = oracle_phase_through_host(files, ["--ont", "--indels"], str(tmp_path / "b"))
    finally:
        os.environ.pop("LPS_READ_SPLIT", None)
    for name in ("chrA", "chrB"):
        a, b = plain[name], split[name]
        assert a.n_reads == b.n_reads > 300 and a.read_names == b.read_names
        for k in ("ref_start", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "flag", "mapq", "name_rank", "cigar", "seq4", "qual"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (name, k)
    assert open(tmp_path / "a" / "out.vcf").read() == open(tmp_path / "b" / "out.vcf").read()


@needs_host
def test_pack_round_trips_the_synthetic_batch(tmp_path_factory, tmp_path):
    """What htslib decodes and the host packs is the batch the generator made (region filter chr:1-lastSNP applied)."""
    files = dataset(tmp_path_factory, "plain")
    seen = oracle_phase_through_host(files, ["--ont", "--indels"], str(tmp_path / "own"))
    src = hc.synth.Contig(seed=201, contig_len=260_000, indel_frac=0.12, depth=18.0, mean_len=9_000.0)
    c = seen["chrA"]
    last = int(c.var_pos[-1])
    keep = np.nonzero(src.ref_start < last)[0]          # 1-based region chr:1-last = 0-based [0, last)
    assert c.n_reads == len(keep)
    assert np.array_equal(c.ref_start, src.ref_start[keep]) and np.array_equal(c.flag, src.flag[keep]) and np.array_equal(c.mapq, src.mapq[keep])
    assert np.array_equal(c.n_cigar, src.n_cigar[keep]) and np.array_equal(c.l_qseq, src.l_qseq[keep])
    for k in (0, len(keep) // 2, len(keep) - 1):
        r = int(keep[k])
        a, n = int(src.cigar_off[r]), int(src.n_cigar[r])
        assert np.array_equal(c.cigar[int(c.cigar_off[k]):int(c.cigar_off[k]) + n], src.cigar[a:a + n])
        lq = int(src.l_qseq[r])
        assert np.array_equal(c.qual[int(c.qual_off[k]):int(c.qual_off[k]) + lq], np.minimum(src.qual[int(src.qual_off[r]):int(src.qual_off[r]) + lq], 93))
        nb = lq // 2
        assert np.array_equal(c.seq4[int(c.seq_off[k]):int(c.seq_off[k]) + nb], src.seq4[int(src.seq_off[r]):int(src.seq_off[r]) + nb])
        assert c.read_names[k] == src.name(r)
    # name ranks: the order of std::string names, equal names share a rank
    order = sorted(range(c.n_reads), key=lambda k: c.read_names[k].encode())
    for x, y in zip(order, order[1:]):
        same = c.read_names[x] == c.read_names[y]
        assert (c.name_rank[x] == c.name_rank[y]) if same else (c.name_rank[x] < c.name_rank[y])
    # the reference string ends 5 bases after the last variant (FastaParser, ParsingBam.cpp:46)
    assert c.ref == src.ref[:last + 5 + 1] or c.ref == src.ref[:last + 5]


def tag_args(files, vcf, extra):
    return ["haplotag", "-s", vcf, "-b", files["bam"], "-r", files["fasta"], "-o", "tagged", "-t", "2"] + extra


def oracle_tag_through_host(files, vcf, extra, cwd, chunk=None):
    lib = hc.host_lib()
    os.makedirs(cwd, exist_ok=True)
    old = os.getcwd()
    os.chdir(cwd)
    if chunk:
        os.environ["LPS_TAG_CHUNK"] = str(chunk)
    try:
        h = C.c_void_p()
        n, av = hc.argv(tag_args(files, vcf, extra))
        assert lib.lpsh_tag_open(n, av, C.byref(h)) == 0, lib.lpsh_last_error()
        tp = ffi.LpsTagParams()
        assert lib.lpsh_tag_params(h, C.byref(tp)) == 0
        assert lib.lpsh_tag_begin(h) == 0, lib.lpsh_last_error()
        chunks = 0
        for i in range(lib.lpsh_tag_n_contigs(h)):
            while True:
                pk = hc.LpshPacked()
                got = lib.lpsh_tag_pack(h, i, C.byref(pk))
                assert got >= 0, lib.lpsh_last_error()
                if got == 0:
                    break
                chunks += 1
                c = hc.packed_contig(pk)
                r = ffi.LpsTagResult()
                r.n_reads = c.n_reads
                if c.n_var == 0:                      # the host's own dispatch for a contig without variants
                    cat = np.where(c.mapq < tp.mapping_quality, 1, np.where(c.flag & 4, 2, np.where(c.flag & 0x100, 3, np.where(
                        (c.flag & 0x800 != 0) & (tp.tag_supplementary == 0), 4, 5)))).astype(np.uint8)
                    keep = [cat, np.zeros(c.n_reads, np.int8), np.zeros(c.n_reads, np.int32)]
                    r.category, r.hp = ffi.ptr(keep[0], ffi.u8p), ffi.ptr(keep[1], ffi.i8p)
                    r.ps = r.pq = r.h1 = r.h2 = ffi.ptr(keep[2], ffi.i32p)
                else:
                    orc = po.OracleTag(c, tp)
                    assert orc.rc == 0
                    keep = [np.ascontiguousarray(x) for x in (orc.category, orc.hp, orc.ps, orc.pq, orc.h1, orc.h2, orc.call_off, orc.calls)]
                    r.category, r.hp, r.ps, r.pq = ffi.ptr(keep[0], ffi.u8p), ffi.ptr(keep[1], ffi.i8p), ffi.ptr(keep[2], ffi.i32p), ffi.ptr(keep[3], ffi.i32p)
                    r.h1, r.h2, r.call_off = ffi.ptr(keep[4], ffi.i32p), ffi.ptr(keep[5], ffi.i32p), ffi.ptr(keep[6], ffi.u64p)
                    r.n_calls = len(keep[7])
                    r.calls = keep[7].ctypes.data_as(C.POINTER(ffi.LpsCall))
                assert lib.lpsh_tag_emit(h, i, C.byref(r)) == 0, lib.lpsh_last_error()
        assert lib.lpsh_tag_end(h) == 0
        lib.lpsh_tag_close(h)
        return chunks
    finally:
        os.environ.pop("LPS_TAG_CHUNK", None)
        os.chdir(old)


TAG_VARIANTS = [[], ["--log"], ["--log", "--tagSupplementary", "-q", "20", "-p", "0.75"], ["--log", "--region", "chrB:20000-120000"]]


@needs_host
@needs_ref
@pytest.mark.parametrize("extra", TAG_VARIANTS)
def test_haplotag_host_files_match_reference(tmp_path_factory, tmp_path, extra):
    files = dataset(tmp_path_factory, "plain")
    key = "phased_vcf"
    if key not in files:                                  # the phased VCF both programs tag with: the reference's own phase output
        d = os.path.join(files["dir"], "phase_ref")
        run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
        files[key] = os.path.join(d, "out.vcf")
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, files[key], extra))
    chunks = oracle_tag_through_host(files, files[key], extra, str(tmp_path / "own"), chunk=100 if "--region" in extra else 700)
    assert chunks >= 3
    ref, own = hc.bam_payload(str(tmp_path / "ref" / "tagged.bam")), hc.bam_payload(str(tmp_path / "own" / "tagged.bam"))
    assert len(own) > 100_000 and b"HP" in own
    assert own == ref, "tagged BAM differs from the reference's (uncompressed byte stream)"
    if "--log" in extra:
        assert open(tmp_path / "own" / "tagged.out").read() == open(tmp_path / "ref" / "tagged.out").read()


def oracle_tag_pipelined(files, vcf, extra, cwd, chunk):
    """lpsh_tag_run_with: the host's own reader-thread / writer pipeline with the ORACLE as the judge of every chunk."""
    lib = hc.host_lib()
    os.makedirs(cwd, exist_ok=True)
    old = os.getcwd()
    os.chdir(cwd)
    os.environ["LPS_TAG_CHUNK"] = str(chunk)
    try:
        h = C.c_void_p()
        n, av = hc.argv(tag_args(files, vcf, extra))
        assert lib.lpsh_tag_open(n, av, C.byref(h)) == 0, lib.lpsh_last_error()
        tp = ffi.LpsTagParams()
        lib.lpsh_tag_params(h, C.byref(tp))
        state = dict(keep=None, chunks=0, contigs=set())

        def judge(user, contig, pk, want_calls, out):
            try:
                c = hc.packed_contig(pk.contents)
                orc = po.OracleTag(c, tp)
                keep = [np.ascontiguousarray(x) for x in (orc.category, orc.hp, orc.ps, orc.pq, orc.h1, orc.h2, orc.call_off, orc.calls)]
                r = out.contents
                r.n_reads = c.n_reads
                r.category, r.hp, r.ps, r.pq = ffi.ptr(keep[0], ffi.u8p), ffi.ptr(keep[1], ffi.i8p), ffi.ptr(keep[2], ffi.i32p), ffi.ptr(keep[3], ffi.i32p)
                r.h1, r.h2 = ffi.ptr(keep[4], ffi.i32p), ffi.ptr(keep[5], ffi.i32p)
                if want_calls:
                    r.call_off, r.n_calls, r.calls = ffi.ptr(keep[6], ffi.u64p), len(keep[7]), keep[7].ctypes.data_as(C.POINTER(ffi.LpsCall))
                state["keep"] = keep            # alive until the next chunk is judged
                state["chunks"] += 1
                state["contigs"].add(contig)
                return 0 if orc.rc == 0 else -1
            except Exception as e:  # noqa: BLE001
                print("judge failed:", e)
                return -1
        cb = hc.TAG_JUDGE_FN(judge)
        rc = lib.lpsh_tag_run_with(h, cb, None)
        lib.lpsh_tag_close(h)
        assert rc == 0, lib.lpsh_last_error()
        return state
    finally:
        os.environ.pop("LPS_TAG_CHUNK", None)
        os.chdir(old)


@needs_host
@needs_ref
@pytest.mark.parametrize("chunk", [250, 100000])
def test_haplotag_pipelined_run_matches_reference(tmp_path_factory, tmp_path, chunk):
    files = dataset(tmp_path_factory, "plain")
    if "phased_vcf" not in files:
        d = os.path.join(files["dir"], "phase_ref")
        run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
        files["phased_vcf"] = os.path.join(d, "out.vcf")
    extra = ["--log", "--tagSupplementary"]
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, files["phased_vcf"], extra))
    st = oracle_tag_pipelined(files, files["phased_vcf"], extra, str(tmp_path / "own"), chunk)
    assert st["contigs"] == {0, 2} and (st["chunks"] >= 4 if chunk == 250 else st["chunks"] == 2)   # chrEmpty never reaches the judge
    assert hc.bam_payload(str(tmp_path / "own" / "tagged.bam")) == hc.bam_payload(str(tmp_path / "ref" / "tagged.bam"))
    assert open(tmp_path / "own" / "tagged.out").read() == open(tmp_path / "ref" / "tagged.out").read()


DEFLATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_uint64, C.c_uint32, C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint64))


def host_compiled_deflater(calls):
    """The hook a box without a GPU hands to lpsh_set_deflater: the member encoder of k_bgzf_deflate compiled for the host
    (lps_bgzf_deflate_block_host), member by member - what lps_bgzf_deflate's kernel writes, byte for byte (tests/test_bgzf_deflate.py)."""
    lps = ffi.load_library()

    def deflate(user, src, n, block_bytes, dst, cap, out_len):
        at, n_out = 0, 0
        slot = (C.c_uint8 * 65312)()
        got = C.c_uint32(0)
        while at < n:
            m = min(block_bytes, n - at)
            piece = C.cast(C.addressof(src.contents) + at, ffi.u8p)
            if lps.lps_bgzf_deflate_block_host(piece, m, C.cast(slot, ffi.u8p), 65312, C.byref(got)) != 0 or n_out + got.value > cap:
                return -1
            C.memmove(C.addressof(dst.contents) + n_out, slot, got.value)
            at += m
            n_out += got.value
        out_len[0] = n_out
        calls.append((int(n), n_out))
        return 0
    return DEFLATE_FN(deflate)


@needs_host
@needs_ref
@pytest.mark.parametrize("chunk", [300, 100000])
def test_device_bam_writer_writes_the_records_htslib_writes(tmp_path_factory, tmp_path, chunk, monkeypatch):
    """LPS_GPU_DEFLATE=1: records serialised by the host's own writer and deflated in batches (here by the host-compiled member
    encoder through lpsh_set_deflater; on the GPU box by lps_bgzf_deflate): the uncompressed stream - header, every record, the
    EOF marker's empty member - is what the reference binary's htslib wrote, and htslib reads the file back."""
    files = dataset(tmp_path_factory, "plain")
    vcf = files.get("phased_vcf")
    if not vcf:
        d = os.path.join(files["dir"], "phase_ref")
        run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
        vcf = files["phased_vcf"] = os.path.join(d, "out.vcf")
    extra = ["--tagSupplementary"]
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, vcf, extra))
    lib = hc.host_lib()
    lib.lpsh_set_deflater.argtypes = [DEFLATE_FN, C.c_void_p]
    calls = []
    cb = host_compiled_deflater(calls)
    lib.lpsh_set_deflater(cb, None)
    monkeypatch.setenv("LPS_GPU_DEFLATE", "1")
    if chunk == 300:
        monkeypatch.setenv("LPS_DEFLATE_BATCH", str(8 * 65280))          # many batches: the flusher thread works beside the tagging loop
    try:
        oracle_tag_pipelined(files, vcf, extra, str(tmp_path / "own"), chunk)
    finally:
        lib.lpsh_set_deflater(C.cast(None, DEFLATE_FN), None)
    own_path, ref_path = str(tmp_path / "own" / "tagged.bam"), str(tmp_path / "ref" / "tagged.bam")
    assert calls and sum(c[0] for c in calls) == len(hc.bam_payload(ref_path))
    # the header travels alone (its own members, as bam_hdr_write's flush leaves it), then whole members until the last batch
    assert (len(calls) > 10 and all(c[0] % 65280 == 0 for c in calls[1:-1])) if chunk == 300 else len(calls) == 2
    first_isize = lambda path: int.from_bytes(open(path, "rb").read()[:65536][int.from_bytes(open(path, "rb").read()[16:18], "little") + 1 - 4:][:4], "little")  # noqa: E731
    assert first_isize(own_path) == first_isize(ref_path) == calls[0][0], "the BAM header does not sit in a member of its own" 
    assert hc.bam_payload(own_path) == hc.bam_payload(ref_path)
    raw = open(own_path, "rb").read()
    assert raw.endswith(bytes([0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0]))
    # htslib's own reader (sam_read1 with a thread pool) walks the file to its end
    lib.lpsh_decode_only.restype = C.c_int64
    lib.lpsh_decode_only.argtypes = [C.c_char_p, C.c_int]
    n_ref = lib.lpsh_decode_only(ref_path.encode(), 2)
    assert n_ref > 0 and lib.lpsh_decode_only(own_path.encode(), 2) == n_ref


@needs_host
@needs_ref
def test_device_bam_writer_long_cigar_record(tmp_path, monkeypatch):
    """A record with more than 65 535 CIGAR operations: BAM keeps <l_qseq>S<rlen>N in the CIGAR field and the real operations in a
    CG:B,I tag (SAM spec 4.2.2; bam_write1, htslib/sam.c:836-860).  The device writer must lay it out the same way."""
    from tests import handmade
    rng = np.random.default_rng(77)
    ref = "".join(rng.choice(list("ACGT"), 50_000))
    variants = []
    for p in range(500, 49_000, 700):
        alt = "ACGT"[("ACGT".index(ref[p]) + 1) % 4]
        variants.append((p, ref[p], alt, p // 700 % 2))
    R = handmade.read_from_ref
    big = "".join("1M1I" for _ in range(33_000)) + "500M"                        # 66 001 operations, 33 500 reference bases
    reads = [R(ref, 100, "3000M", "r_a"), R(ref, 200, big, "r_long"), R(ref, 300, "20S4000M10S", "r_b"), R(ref, 30_000, "6S9000M", "r_c"),
             R(ref, 40_000, "8000M7S", "r_d")]
    c = handmade.ManualContig(ref, variants, reads)
    assert int(c.n_cigar.max()) > 65535
    os.makedirs(tmp_path / "data")
    files = hc.write_dataset(str(tmp_path / "data"), [("chrL", c, True)])
    d = str(tmp_path / "phase_ref")
    run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont"]))
    vcf = os.path.join(d, "out.vcf")
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, vcf, []))
    lib = hc.host_lib()
    lib.lpsh_set_deflater.argtypes = [DEFLATE_FN, C.c_void_p]
    calls = []
    cb = host_compiled_deflater(calls)
    lib.lpsh_set_deflater(cb, None)
    monkeypatch.setenv("LPS_GPU_DEFLATE", "1")
    try:
        oracle_tag_pipelined(files, vcf, [], str(tmp_path / "own"), 100000)
    finally:
        lib.lpsh_set_deflater(C.cast(None, DEFLATE_FN), None)
    own, want = hc.bam_payload(str(tmp_path / "own" / "tagged.bam")), hc.bam_payload(str(tmp_path / "ref" / "tagged.bam"))
    assert b"CGBI" in want and own == want


@needs_host
@needs_ref
def test_bam_level_changes_the_file_not_the_records(tmp_path_factory, tmp_path):
    """LPS_BAM_LEVEL=1: a faster deflate level for the tagged BAM (the writer is what the pass waits for); same uncompressed stream."""
    files = dataset(tmp_path_factory, "plain")
    vcf = files.get("phased_vcf")
    if not vcf:
        d = os.path.join(files["dir"], "phase_ref")
        run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
        vcf = files["phased_vcf"] = os.path.join(d, "out.vcf")
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, vcf, []))
    os.environ["LPS_BAM_LEVEL"] = "1"
    try:
        oracle_tag_pipelined(files, vcf, [], str(tmp_path / "own"), 400)
    finally:
        os.environ.pop("LPS_BAM_LEVEL", None)
    assert hc.bam_payload(str(tmp_path / "own" / "tagged.bam")) == hc.bam_payload(str(tmp_path / "ref" / "tagged.bam"))
    assert os.path.getsize(tmp_path / "own" / "tagged.bam") > os.path.getsize(tmp_path / "ref" / "tagged.bam")


@needs_host
@needs_ref
def test_haplotag_cram_output(tmp_path_factory, tmp_path):
    """--cram: the tagged alignments as CRAM (hts_open "wc" with the FASTA, HaplotagParsingBam.cpp:56-60); decoded back to SAM text,
    the file equals the reference's."""
    files = dataset(tmp_path_factory, "plain")
    if "phased_vcf" not in files:
        d = os.path.join(files["dir"], "phase_ref")
        run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
        files["phased_vcf"] = os.path.join(d, "out.vcf")
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, files["phased_vcf"], ["--cram"]))
    oracle_tag_pipelined(files, files["phased_vcf"], ["--cram"], str(tmp_path / "own"), 400)
    lib = hc.host_lib()
    lib.lpsh_to_sam.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    for side in ("ref", "own"):
        assert lib.lpsh_to_sam(str(tmp_path / side / "tagged.cram").encode(), files["fasta"].encode(), str(tmp_path / side / "tagged.sam").encode()) == 0
    own = open(tmp_path / "own" / "tagged.sam").read()
    assert own == open(tmp_path / "ref" / "tagged.sam").read() and "\tHP:i:" in own and own.count("\n") > 900


@needs_host
@needs_ref
@pytest.mark.parametrize("extra", [["--log", "--tagSupplementary"], ["--log", "--region", "chrB:20000-120000"], ["--region", "chrA"]])
def test_haplotag_ordered_slice_readers(tmp_path_factory, tmp_path, extra):
    """LPS_TAG_READERS=3: position slices of every contig read by three threads with their own file handles, handled in slice order;
    tagged BAM and .out are the serial loop's, i.e. the reference's."""
    files = dataset(tmp_path_factory, "plain")
    if "phased_vcf" not in files:
        d = os.path.join(files["dir"], "phase_ref")
        run_in(d, [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
        files["phased_vcf"] = os.path.join(d, "out.vcf")
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, files["phased_vcf"], extra))
    os.environ.update(LPS_TAG_READERS="3", LPS_TAG_SLICE_BP="30000")
    try:
        st = oracle_tag_pipelined(files, files["phased_vcf"], extra, str(tmp_path / "own"), 8192)
    finally:
        os.environ.pop("LPS_TAG_READERS", None)
        os.environ.pop("LPS_TAG_SLICE_BP", None)
    assert st["chunks"] >= 4
    assert hc.bam_payload(str(tmp_path / "own" / "tagged.bam")) == hc.bam_payload(str(tmp_path / "ref" / "tagged.bam"))
    if "--log" in extra:
        assert open(tmp_path / "own" / "tagged.out").read() == open(tmp_path / "ref" / "tagged.out").read()


@needs_host
def test_haplotag_pipelined_run_stops_on_judge_failure(tmp_path_factory, tmp_path):
    files = dataset(tmp_path_factory, "plain")
    lib = hc.host_lib()
    old = os.getcwd()
    os.chdir(str(tmp_path))
    os.environ["LPS_TAG_CHUNK"] = "50"
    try:
        h = C.c_void_p()
        n, av = hc.argv(tag_args(files, files.get("phased_vcf", files["vcf"]), []))
        assert lib.lpsh_tag_open(n, av, C.byref(h)) == 0
        calls = []
        cb = hc.TAG_JUDGE_FN(lambda user, contig, pk, want, out: calls.append(contig) or (-1 if len(calls) == 3 else _zero_verdicts(pk, out, calls)))
        assert lib.lpsh_tag_run_with(h, cb, None) != 0 and len(calls) == 3
        assert b"judge failed" in lib.lpsh_last_error()
        lib.lpsh_tag_close(h)
    finally:
        os.environ.pop("LPS_TAG_CHUNK", None)
        os.chdir(old)


_zero_keep = []


def _zero_verdicts(pk, out, calls):
    n = pk.contents.batch.n_reads
    cat, z8, z = np.ones(n, np.uint8), np.zeros(n, np.int8), np.zeros(n, np.int32)
    _zero_keep[:] = [cat, z8, z]
    r = out.contents
    r.n_reads, r.category, r.hp = n, ffi.ptr(cat, ffi.u8p), ffi.ptr(z8, ffi.i8p)
    r.ps = r.pq = r.h1 = r.h2 = ffi.ptr(z, ffi.i32p)
    return 0


@needs_host
@needs_ref
def test_haplotag_vcf_loader_quirks(tmp_path):
    """The phased VCF as other tools write it: PS before GT, GT last, a second ALT allele (kept unless the sample text holds a '2',
    HaplotagVcfParser.cpp:283-286), MNP records.  Same tags and the same .out as the reference."""
    import random
    d = str(tmp_path)
    a = hc.synth.Contig(seed=311, contig_len=200_000, indel_frac=0.15, depth=14.0, mean_len=7_000.0)
    files = hc.write_dataset(d, [("chrA", a, True)], fast_bam=True)
    run_in(os.path.join(d, "p"), [hc.REF_BIN] + phase_args(files, ["--ont", "--indels"]))
    rng, ref, out, touched = random.Random(3), a.ref.decode(), [], 0
    for ln in open(os.path.join(d, "p", "out.vcf")).read().split("\n"):
        if ln and not ln.startswith("#"):
            t = ln.split("\t")
            fmt = t[8].split(":")
            kv = dict(zip(fmt, t[9].split(":")))
            u = rng.random()
            order = (["PS", "GT"] + [k for k in fmt if k not in ("PS", "GT")]) if (u < 0.2 and "PS" in kv) else \
                ([k for k in fmt if k != "GT"] + ["GT"]) if u < 0.3 else fmt
            t[8], t[9] = ":".join(order), ":".join(kv[k] for k in order)
            if u > 0.9 and len(t[3]) == 1 and len(t[4]) == 1:
                t[4] += "," + rng.choice("ACGT")
            if 0.85 < u < 0.9 and len(t[3]) == 1 and len(t[4]) == 1 and int(t[1]) + 2 < len(ref):
                p = int(t[1])
                t[3], t[4] = ref[p - 1:p + 1].upper(), t[4] + rng.choice("ACGT")
            touched += order != fmt or "," in t[4] or len(t[3]) == 2
            ln = "\t".join(t)
        out.append(ln)
    assert touched > 40
    vcf = os.path.join(d, "odd.vcf")
    open(vcf, "w").write("\n".join(out))
    run_in(os.path.join(d, "ref"), [hc.REF_BIN] + tag_args(files, vcf, ["--log"]))
    oracle_tag_pipelined(files, vcf, ["--log"], os.path.join(d, "own"), 500)
    assert hc.bam_payload(os.path.join(d, "own", "tagged.bam")) == hc.bam_payload(os.path.join(d, "ref", "tagged.bam"))
    assert open(os.path.join(d, "own", "tagged.out")).read() == open(os.path.join(d, "ref", "tagged.out")).read()


@needs_host
@needs_ref
def test_haplotag_string_phase_sets(tmp_path_factory, tmp_path):
    """##FORMAT=<ID=PS,...Type=String>: phase-set names are indexed in order of first appearance, from 0 (HaplotagVcfParser.cpp:316-320)."""
    files = dataset(tmp_path_factory, "plain")
    out, i = [], 0
    for ln in open(files["vcf"]).read().split("\n"):
        if ln.startswith("#CHROM"):
            out.append('##FORMAT=<ID=PS,Number=1,Type=String,Description="ps">')
        if ln and not ln.startswith("#"):
            t = ln.split("\t")
            if len(t[3]) == 1 and "," not in t[4]:
                t[8], t[9] = "GT:PS", ("0|1" if i % 2 else "1|0") + ":blk%d" % (i // 40)
                i += 1
                ln = "\t".join(t)
        out.append(ln)
    vcf = str(tmp_path / "string_ps.vcf")
    open(vcf, "w").write("\n".join(out))
    run_in(str(tmp_path / "ref"), [hc.REF_BIN] + tag_args(files, vcf, ["--log"]))
    oracle_tag_through_host(files, vcf, ["--log"], str(tmp_path / "own"), chunk=5000)
    assert hc.bam_payload(str(tmp_path / "own" / "tagged.bam")) == hc.bam_payload(str(tmp_path / "ref" / "tagged.bam"))
    log = open(tmp_path / "own" / "tagged.out").read()
    assert log == open(tmp_path / "ref" / "tagged.out").read()
    seen = {ln.split("\t")[5] for ln in log.split("\n") if ln and ln[0] != "#"} - {"."}
    assert "0" in seen and len(seen) >= 3


@needs_host
def test_host_rejects_bad_options(tmp_path, capfd):
    lib = hc.host_lib()
    h = C.c_void_p()
    n, av = hc.argv(["phase", "-s", "/nonexistent.vcf", "-b", "x.bam", "-r", "/nonexistent.fa"])
    assert lib.lpsh_phase_open(n, av, C.byref(h)) == 1 and not h.value
    err = capfd.readouterr().err
    assert "--ont or --pb" in err and "not exist" in err
    n, av = hc.argv(["haplotag", "-b", "x.bam", "-p", "1.5"])
    assert lib.lpsh_tag_open(n, av, C.byref(h)) == 1 and not h.value
    assert "missing SNP file" in capfd.readouterr().err


@needs_host
def test_host_binary_needs_a_gpu(tmp_path_factory, tmp_path):
    """No CPU fallback: without a CUDA device the binary stops with an error instead of computing anything on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    files = dataset(tmp_path_factory, "plain")
    p = subprocess.run([hc.HOST_BIN] + phase_args(files, ["--ont"]), cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and "no usable CUDA device" in p.stderr and not os.path.exists(tmp_path / "out.vcf")
    p = subprocess.run([hc.HOST_BIN] + tag_args(files, files["vcf"], []), cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and "no usable CUDA device" in p.stderr


# ---- the real thing: the binary, device and all, against the reference binary ---------------------------------------------
@pytest.mark.gpu
@needs_host
@needs_ref
def test_gpu_cli_phase_and_haplotag_match_reference(tmp_path_factory, tmp_path):
    files = dataset(tmp_path_factory, "plain")
    for k, extra in enumerate([["--ont", "--indels"], ["--pb"]]):
        run_in(str(tmp_path / f"ref{k}"), [hc.REF_BIN] + phase_args(files, extra))
        run_in(str(tmp_path / f"own{k}"), [hc.HOST_BIN] + phase_args(files, extra))
        ref, own = open(tmp_path / f"ref{k}" / "out.vcf").read(), open(tmp_path / f"own{k}" / "out.vcf").read()
        assert hc.strip_commandline(own) == hc.strip_commandline(ref), extra
    phased = str(tmp_path / "own0" / "out.vcf")
    for k, extra in enumerate([["--log"], ["--log", "--tagSupplementary", "-q", "20", "-p", "0.75"]]):
        run_in(str(tmp_path / f"tref{k}"), [hc.REF_BIN] + tag_args(files, phased, extra))
        env_chunk = dict(os.environ, LPS_TAG_CHUNK="900") if k else None
        os.makedirs(tmp_path / f"town{k}", exist_ok=True)
        p = subprocess.run([hc.HOST_BIN] + tag_args(files, phased, extra), cwd=str(tmp_path / f"town{k}"), env=env_chunk,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        assert hc.bam_payload(str(tmp_path / f"town{k}" / "tagged.bam")) == hc.bam_payload(str(tmp_path / f"tref{k}" / "tagged.bam")), extra
        assert open(tmp_path / f"town{k}" / "tagged.out").read() == open(tmp_path / f"tref{k}" / "tagged.out").read(), extra
