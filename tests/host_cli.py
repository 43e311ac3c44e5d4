"""Helpers of the C++ host tests (tests/test_host_cli.py): a small multi-contig data set on disk (FASTA, VCF, BAM + BAI written
by the reference's own htslib through oracle/_ref/mkbam), ctypes bindings of liblps_host.so (longphase-s_b200/host/lps_host.h),
and views of a packed contig that the oracle accepts.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import gzip
import importlib
import os
import subprocess

import numpy as np

import __graft_entry__ as entry

entry.load_package()
synth = importlib.import_module("longphase_s_b200.synth")
ffi = importlib.import_module("longphase_s_b200._ffi")

ROOT = entry.ROOT
HOST_LIB = os.path.join(entry.PKG_DIR, "liblps_host.so")
HOST_BIN = os.path.join(entry.PKG_DIR, "longphase-s-b200")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "longphase-s")
MKBAM = os.path.join(ROOT, "oracle", "_ref", "mkbam")

OPS = "MIDNSHP=XB"
NT16 = "=ACMGRSVTWYHKDBN"


class LpshPacked(C.Structure):
    _fields_ = [("variants", ffi.LpsVariants), ("batch", ffi.LpsReadBatch), ("ref", C.POINTER(C.c_char)), ("ref_len", C.c_int64),
                ("names", C.POINTER(C.c_char)), ("name_off", ffi.u64p)]


TAG_JUDGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(LpshPacked), C.c_int, C.POINTER(ffi.LpsTagResult))

_lib = None


def host_lib():
    global _lib
    if _lib is None:
        lib = C.CDLL(HOST_LIB)
        vp = C.c_void_p
        lib.lpsh_phase_open.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp)]
        lib.lpsh_phase_n_contigs.argtypes = [vp]
        lib.lpsh_phase_contig_name.argtypes = [vp, C.c_int]
        lib.lpsh_phase_contig_name.restype = C.c_char_p
        lib.lpsh_phase_last_variant.argtypes = [vp, C.c_int]
        lib.lpsh_phase_params.argtypes = [vp, C.POINTER(ffi.LpsPhaseParams)]
        lib.lpsh_phase_pack.argtypes = [vp, C.c_int, C.POINTER(LpshPacked)]
        lib.lpsh_phase_release.argtypes = [vp, C.c_int]
        lib.lpsh_phase_set_result.argtypes = [vp, C.c_int, C.c_int32, ffi.i32p, ffi.i8p]
        lib.lpsh_phase_write_result.argtypes = [vp]
        lib.lpsh_phase_close.argtypes = [vp]
        lib.lpsh_tag_open.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp)]
        lib.lpsh_tag_n_contigs.argtypes = [vp]
        lib.lpsh_tag_contig_name.argtypes = [vp, C.c_int]
        lib.lpsh_tag_contig_name.restype = C.c_char_p
        lib.lpsh_tag_params.argtypes = [vp, C.POINTER(ffi.LpsTagParams)]
        lib.lpsh_tag_begin.argtypes = [vp]
        lib.lpsh_tag_pack.argtypes = [vp, C.c_int, C.POINTER(LpshPacked)]
        lib.lpsh_tag_emit.argtypes = [vp, C.c_int, C.POINTER(ffi.LpsTagResult)]
        lib.lpsh_tag_end.argtypes = [vp]
        lib.lpsh_tag_run_with.argtypes = [vp, TAG_JUDGE_FN, vp]
        lib.lpsh_tag_close.argtypes = [vp]
        lib.lpsh_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def argv(words):
    arr = (C.c_char_p * (len(words) + 1))()
    for i, w in enumerate(words):
        arr[i] = w.encode()
    return len(words), arr


def packed_contig(pk):
    """A synth.Contig-shaped numpy COPY of an lpsh_packed (so that the oracle and compare helpers take it as they take synthetic contigs)."""
    g = ffi.as_np
    c = synth.Contig.__new__(synth.Contig)
    v, b = pk.variants, pk.batch
    nv, nr = v.n, b.n_reads
    c.n_var, c.n_reads = nv, nr
    c.ref = C.string_at(pk.ref, pk.ref_len)
    c.var_pos = g(v.pos, nv, np.int32)
    c.var_ref0, c.var_alt0 = g(v.ref0, nv, np.uint8), g(v.alt0, nv, np.uint8)
    c.var_ref_len, c.var_alt_len = g(v.ref_len, nv, np.uint16), g(v.alt_len, nv, np.uint16)
    if v.ps:
        c.var_hp1_is_alt, c.var_ps, c.var_gt_kind = g(v.hp1_is_alt, nv, np.uint8), g(v.ps, nv, np.int32), g(v.gt_kind, nv, np.uint8)
    else:
        c.var_hp1_is_alt = np.zeros(nv, np.uint8)
    c.ref_start, c.l_qseq = g(b.ref_start, nr, np.int32), g(b.l_qseq, nr, np.int32)
    c.n_cigar = g(b.n_cigar, nr, np.uint32)
    c.cigar_off, c.seq_off, c.qual_off = g(b.cigar_off, nr, np.uint64), g(b.seq_off, nr, np.uint64), g(b.qual_off, nr, np.uint64)
    c.flag, c.mapq, c.name_rank = g(b.flag, nr, np.uint16), g(b.mapq, nr, np.uint8), g(b.name_rank, nr, np.int32)
    c.cigar, c.seq4, c.qual = g(b.cigar, b.cigar_len, np.uint32), g(b.seq4, b.seq_bytes, np.uint8), g(b.qual, b.qual_bytes, np.uint8)
    off = g(pk.name_off, nr, np.uint64)
    c.read_names = [C.string_at(C.addressof(pk.names.contents) + int(o)).decode() for o in off] if nr else []
    return c


# ---- data set on disk --------------------------------------------------------------------------------------------------
def _sam_records(contig, chrom, f):
    for r in range(contig.n_reads):
        co, nc = int(contig.cigar_off[r]), int(contig.n_cigar[r])
        cigar = "".join("%d%s" % (w >> 4, OPS[w & 15]) for w in contig.cigar[co:co + nc].tolist())
        lq = int(contig.l_qseq[r])
        so, qo = int(contig.seq_off[r]), int(contig.qual_off[r])
        packed = contig.seq4[so:so + (lq + 1) // 2]
        codes = np.stack([packed >> 4, packed & 15], 1).reshape(-1)[:lq]
        seq = "".join(NT16[c] for c in codes.tolist()) or "*"
        qual = "".join(chr(33 + min(int(q), 93)) for q in contig.qual[qo:qo + lq].tolist()) or "*"
        aux = "\tHP:i:2\tPS:i:7\tXX:Z:keep" if r % 11 == 0 else ("\tNM:i:3" if r % 3 == 0 else "")   # old tags must be replaced, others kept
        f.write("\t".join([contig.name(r), str(int(contig.flag[r])), chrom, str(int(contig.ref_start[r]) + 1), str(int(contig.mapq[r])),
                           cigar or "*", "*", "0", "0", seq, qual]) + aux + "\n")


GT_FORMS = ["0/1", "1/0", "0|1", "1|0"]


def write_bam_fast(bam, contigs, threads=8):
    """The generator's batches straight to BAM + BAI through the host library's htslib writer (no SAM text, no aux tags)."""
    lib = C.CDLL(HOST_LIB)
    lib.lpsh_bamw_open.restype = C.c_void_p
    lib.lpsh_bamw_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int]
    lib.lpsh_bamw_append.argtypes = [C.c_void_p, C.c_int, C.POINTER(ffi.LpsReadBatch), C.c_char_p, C.c_int]
    lib.lpsh_bamw_close.argtypes = [C.c_void_p]
    n = len(contigs)
    names = (C.c_char_p * n)(*[name.encode() for name, _, _ in contigs])
    lens = (C.c_int64 * n)(*[len(c.ref) for _, c, _ in contigs])
    w = lib.lpsh_bamw_open(bam.encode(), n, names, lens, threads)
    assert w, "lpsh_bamw_open failed"
    for tid, (_, c, _) in enumerate(contigs):
        bs = c.batch_struct()
        assert lib.lpsh_bamw_append(w, tid, C.byref(bs), c.names, c.NAME_STRIDE) == 0
    assert lib.lpsh_bamw_close(w) == 0


def write_dataset(d, contigs, with_ps=False, gz=False, fast_bam=False):
    """contigs: list of (name, synth.Contig, has_variants); has_variants False = a contig with reads but no VCF record.
    Returns dict(vcf, bam, fasta)."""
    fasta, sam, bam = os.path.join(d, "ref.fa"), os.path.join(d, "reads.sam"), os.path.join(d, "reads.bam")
    vcf = os.path.join(d, "calls.vcf")
    with open(fasta, "w") as f:
        for name, c, _ in contigs:
            f.write(">%s\n" % name)
            s = c.ref.decode()
            f.write("\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    if fast_bam:
        write_bam_fast(bam, contigs)
    else:
        with open(sam, "w") as f:
            f.write("@HD\tVN:1.6\tSO:coordinate\n")
            for name, c, _ in contigs:
                f.write("@SQ\tSN:%s\tLN:%d\n" % (name, len(c.ref)))
            f.write("@PG\tID:synth\tPN:synth\n")
            for name, c, _ in contigs:
                _sam_records(c, name, f)
        subprocess.check_call([MKBAM, sam, bam], stderr=subprocess.DEVNULL)
        os.remove(sam)
    rng = np.random.default_rng(5)
    with open(vcf, "w") as f:
        f.write("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n")
        for name, c, _ in contigs:
            f.write("##contig=<ID=%s,length=%d>\n" % (name, len(c.ref)))
        f.write("##INFO=<ID=DP,Number=1,Type=Integer,Description=\"depth\">\n")
        f.write("##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"q\">\n")
        f.write("##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n")
        if with_ps:
            f.write("##FORMAT=<ID=PS,Number=1,Type=Integer,Description=\"Phase set identifier\">\n")
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n")
        for name, c, has_variants in contigs:
            if not has_variants:
                continue
            taken = set(c.var_pos.tolist())
            for i in range(c.n_var):
                pos = int(c.var_pos[i])
                ref, alt = c.variant_strings(i)
                gt = GT_FORMS[i % 4]
                qual = "%d" % (5 + (i * 7) % 60) if i % 9 else "."
                if with_ps and i % 3 == 0:
                    fmt, smp = "GT:PS:DP", "%s:%d:%d" % (gt, 1000 + i // 10, 20 + i % 9)
                elif with_ps and i % 3 == 1:
                    fmt, smp = "GT:DP:PS", "%s:%d:%d" % (gt, 20 + i % 9, 1000 + i // 10)
                else:
                    fmt, smp = "GT:GQ:DP", "%s:%d:%d" % (gt, 30 + i % 5, 20 + i % 9)
                f.write("%s\t%d\t.\t%s\t%s\t%s\tPASS\tDP=%d\t%s\t%s\n" % (name, pos + 1, ref, alt, qual, 20 + i % 9, fmt, smp))
                # decoys between the real variants: homozygous, multi-allelic, and a hom-ref record
                nxt = int(c.var_pos[i + 1]) if i + 1 < c.n_var else pos + 50
                q = pos + 1 + (nxt - pos) // 2
                if nxt - pos > 8 and q not in taken and q + 1 < len(c.ref) and rng.random() < 0.15:
                    r0 = chr(c.ref[q])
                    a0 = "ACGT"[("ACGT".index(r0.upper()) + 1) % 4] if r0.upper() in "ACGT" else "A"
                    kind = i % 3
                    if kind == 0:
                        f.write("%s\t%d\t.\t%s\t%s\t50\tPASS\t.\tGT:GQ:DP\t1/1:40:30\n" % (name, q + 1, r0, a0))
                    elif kind == 1:
                        a1 = "ACGT"[("ACGT".index(a0) + 1) % 4]
                        if a1 == r0.upper():
                            a1 = "ACGT"[("ACGT".index(a1) + 1) % 4]
                        f.write("%s\t%d\t.\t%s\t%s,%s\t50\tPASS\t.\tGT:GQ:DP\t1/2:40:30\n" % (name, q + 1, r0, a0, a1))
                    else:
                        f.write("%s\t%d\t.\t%s\t%s\t50\tPASS\t.\tGT:GQ:DP\t0/0:40:30\n" % (name, q + 1, r0, a0))
    if gz:
        with open(vcf, "rb") as f, gzip.open(vcf + ".gz", "wb") as g:
            g.write(f.read())
        os.remove(vcf)
        vcf += ".gz"
    return dict(vcf=vcf, bam=bam, fasta=fasta)


def bam_payload(path):
    """The uncompressed byte stream of a BAM file, BGZF member by member (BSIZE at offset 16 of each member's header;
    gzip.decompress on the whole file is quadratic in the number of members)."""
    import zlib
    with open(path, "rb") as f:
        data = f.read()
    out, pos, view = [], 0, memoryview(data)
    while pos < len(data):
        assert data[pos:pos + 4] == b"\x1f\x8b\x08\x04" and data[pos + 12:pos + 14] == b"BC", "not a BGZF member"
        size = int.from_bytes(data[pos + 16:pos + 18], "little") + 1
        out.append(zlib.decompress(view[pos + 18:pos + size - 8], -15))
        pos += size
    return b"".join(out)


def strip_commandline(text):
    return "\n".join(ln for ln in text.split("\n") if not ln.startswith("##commandline="))


# ---- tumor / normal pair on disk (somatic_haplotag) -----------------------------------------------------------------------
def _vcf_header(f, contigs, with_ps):
    f.write("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n")
    for name, ref_len in contigs:
        f.write("##contig=<ID=%s,length=%d>\n" % (name, ref_len))
    f.write("##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n")
    if with_ps:
        f.write("##FORMAT=<ID=PS,Number=1,Type=Integer,Description=\"Phase set identifier\">\n")
    f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n")


def write_somatic_dataset(d, pairs, seed=9):
    """pairs: list of (name, normal synth.Contig, tumor synth.Contig) generated from the same seed (same reference and variants).
    Writes ref.fa, normal.bam, tumor.bam, germline.vcf (unphased heterozygous germline variants, to be phased by `phase`) and
    tumor.vcf (every somatic variant as 0/1, a share of the germline ones as 0/1, 1/1 or phased).  Returns the paths."""
    rng = np.random.default_rng(seed)
    out = {k: os.path.join(d, v) for k, v in dict(fasta="ref.fa", normal_bam="normal.bam", tumor_bam="tumor.bam", germline_vcf="germline.vcf",
                                                  tumor_vcf="tumor.vcf").items()}
    with open(out["fasta"], "w") as f:
        for name, cn, _ in pairs:
            s = cn.ref.decode()
            f.write(">%s\n" % name + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    write_bam_fast(out["normal_bam"], [(name, cn, True) for name, cn, _ in pairs], threads=4)
    write_bam_fast(out["tumor_bam"], [(name, ct, True) for name, _, ct in pairs], threads=4)
    lens = [(name, len(cn.ref)) for name, cn, _ in pairs]
    with open(out["germline_vcf"], "w") as g, open(out["tumor_vcf"], "w") as t:
        _vcf_header(g, lens, False)
        _vcf_header(t, lens, True)
        for name, cn, _ in pairs:
            for i in range(cn.n_var):
                pos = int(cn.var_pos[i]) + 1
                ref, alt = cn.variant_strings(i)
                if cn.var_is_somatic[i]:
                    t.write("%s\t%d\t.\t%s\t%s\t40\tPASS\t.\tGT:DP\t0/1:%d\n" % (name, pos, ref, alt, 30 + i % 7))
                    continue
                g.write("%s\t%d\t.\t%s\t%s\t40\tPASS\t.\tGT:DP\t%s:%d\n" % (name, pos, ref, alt, "0/1" if i % 2 else "1/0", 25 + i % 5))
                u = rng.random()
                if u < 0.05:
                    t.write("%s\t%d\t.\t%s\t%s\t40\tPASS\t.\tGT:DP\t0/1:%d\n" % (name, pos, ref, alt, 30 + i % 7))
                elif u < 0.10:
                    t.write("%s\t%d\t.\t%s\t%s\t40\tPASS\t.\tGT:DP\t1/1:%d\n" % (name, pos, ref, alt, 30 + i % 7))
                elif u < 0.13:
                    t.write("%s\t%d\t.\t%s\t%s\t40\tPASS\t.\tGT:PS:DP\t%s:%d:%d\n" % (name, pos, ref, alt, "0|1" if i % 2 else "1|0", 5000 + i // 30, 30))
    return out


def bam_digest(path):
    """sha256 of a BAM's content that does not depend on the command line: header text without the @PG line `longphase-s` adds
    (its CL field holds the paths of the run), the reference dictionary, and every record byte."""
    import hashlib
    import struct
    raw = bam_payload(path)
    assert raw[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8:8 + l_text].rstrip(b"\0")
    lines = [ln for ln in text.split(b"\n") if not ln.startswith(b"@PG\tID:longphase-s")]
    h = hashlib.sha256(b"\n".join(lines))
    h.update(raw[8 + l_text:])
    return h.hexdigest()


def text_digest(text):
    import hashlib
    return hashlib.sha256(text.encode()).hexdigest()
