"""ctypes wrapper of synth/synth.c: deterministic ONT-like synthetic contigs."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import _ffi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "synth", "synth.c")
LIB = os.path.join(HERE, "synth", "libsynth.so")


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("contig_len", C.c_int64), ("variant_rate", C.c_double),
                ("indel_frac", C.c_double), ("danger_frac", C.c_double), ("homopolymer_frac", C.c_double),
                ("depth", C.c_double), ("mean_len", C.c_double), ("sigma", C.c_double), ("sub_rate", C.c_double),
                ("ins_rate", C.c_double), ("del_rate", C.c_double), ("clip_frac", C.c_double),
                ("supp_frac", C.c_double), ("sec_frac", C.c_double), ("dup_frac", C.c_double),
                ("noseq_frac", C.c_double), ("lowq_mapq_frac", C.c_double), ("zero_mapq_frac", C.c_double),
                ("tumor", C.c_int32), ("somatic_rate", C.c_double), ("purity", C.c_double), ("read_seed", C.c_uint64)]


class SynthOut(C.Structure):
    _fields_ = [("ref_len", C.c_int64), ("ref", C.POINTER(C.c_char)), ("n_var", C.c_int32), ("var_pos", _ffi.i32p),
                ("var_ref0", _ffi.u8p), ("var_alt0", _ffi.u8p), ("var_ref_len", _ffi.u16p), ("var_alt_len", _ffi.u16p),
                ("var_hp1_is_alt", _ffi.u8p), ("var_is_somatic", _ffi.u8p), ("var_str_off", _ffi.u32p), ("var_str", C.POINTER(C.c_char)),
                ("n_reads", C.c_int32), ("ref_start", _ffi.i32p), ("l_qseq", _ffi.i32p), ("n_cigar", _ffi.u32p),
                ("cigar_off", _ffi.u64p), ("seq_off", _ffi.u64p), ("qual_off", _ffi.u64p), ("flag", _ffi.u16p),
                ("mapq", _ffi.u8p), ("name_rank", _ffi.i32p), ("hap", _ffi.u8p), ("cigar", _ffi.u32p),
                ("cigar_len", C.c_uint64), ("seq4", _ffi.u8p), ("seq_bytes", C.c_uint64), ("qual", _ffi.u8p),
                ("qual_bytes", C.c_uint64), ("names", C.POINTER(C.c_char))]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.synth_generate.argtypes = [C.POINTER(SynthParams), C.POINTER(SynthOut)]
        _lib.synth_generate.restype = C.c_int
        _lib.synth_default_params.argtypes = [C.POINTER(SynthParams)]
        _lib.synth_free.argtypes = [C.POINTER(SynthOut)]
    return _lib


class Contig:
    """One synthetic contig: reference, variant table, read batch (all numpy, host memory)."""

    NAME_STRIDE = 40

    def __init__(self, **kw):
        lib = _load()
        p = SynthParams()
        lib.synth_default_params(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown synth parameter {k}")
            setattr(p, k, v)
        self.params = p
        o = SynthOut()
        rc = lib.synth_generate(C.byref(p), C.byref(o))
        if rc != 0:
            raise RuntimeError("synth_generate failed")
        try:
            g = _ffi.as_np
            self.ref = C.string_at(o.ref, o.ref_len)
            nv, nr = o.n_var, o.n_reads
            self.n_var, self.n_reads = nv, nr
            self.var_pos = g(o.var_pos, nv, np.int32)
            self.var_ref0 = g(o.var_ref0, nv, np.uint8)
            self.var_alt0 = g(o.var_alt0, nv, np.uint8)
            self.var_ref_len = g(o.var_ref_len, nv, np.uint16)
            self.var_alt_len = g(o.var_alt_len, nv, np.uint16)
            self.var_hp1_is_alt = g(o.var_hp1_is_alt, nv, np.uint8)
            self.var_is_somatic = g(o.var_is_somatic, nv, np.uint8)
            self.var_str_off = g(o.var_str_off, nv + 1, np.uint32)
            self.var_str = C.string_at(o.var_str, int(self.var_str_off[-1]) if nv else 0)
            self.ref_start = g(o.ref_start, nr, np.int32)
            self.l_qseq = g(o.l_qseq, nr, np.int32)
            self.n_cigar = g(o.n_cigar, nr, np.uint32)
            self.cigar_off = g(o.cigar_off, nr, np.uint64)
            self.seq_off = g(o.seq_off, nr, np.uint64)
            self.qual_off = g(o.qual_off, nr, np.uint64)
            self.flag = g(o.flag, nr, np.uint16)
            self.mapq = g(o.mapq, nr, np.uint8)
            self.name_rank = g(o.name_rank, nr, np.int32)
            self.hap = g(o.hap, nr, np.uint8)
            self.cigar = g(o.cigar, o.cigar_len, np.uint32)
            self.seq4 = g(o.seq4, o.seq_bytes, np.uint8)
            self.qual = g(o.qual, o.qual_bytes, np.uint8)
            self.names = C.string_at(o.names, nr * self.NAME_STRIDE)
        finally:
            lib.synth_free(C.byref(o))

    # ---- views in the C ABI layout -------------------------------------------------------
    def variants_struct(self):
        P = _ffi.ptr
        ps = getattr(self, "var_ps", None)
        gt = getattr(self, "var_gt_kind", None)
        return _ffi.LpsVariants(n=self.n_var, pos=P(self.var_pos, _ffi.i32p), ref0=P(self.var_ref0, _ffi.u8p),
                                alt0=P(self.var_alt0, _ffi.u8p), ref_len=P(self.var_ref_len, _ffi.u16p),
                                alt_len=P(self.var_alt_len, _ffi.u16p), hp1_is_alt=P(self.var_hp1_is_alt, _ffi.u8p),
                                ps=P(ps, _ffi.i32p), gt_kind=P(gt, _ffi.u8p))

    def phased(self, ps, hp1_is_alt=None):
        """The contig as `haplotag` sees it after `phase`: only variants with a phase set (ps != 0) remain, each with
        its PS and the haplotype that carries ALT (GT 1|0 -> hp1_is_alt).  Reads are shared with self."""
        import copy
        keep = np.nonzero(np.asarray(ps) != 0)[0]
        c = copy.copy(self)
        c.n_var = len(keep)
        for k in ("var_pos", "var_ref0", "var_alt0", "var_ref_len", "var_alt_len"):
            setattr(c, k, np.ascontiguousarray(getattr(self, k)[keep]))
        h = self.var_hp1_is_alt if hp1_is_alt is None else np.asarray(hp1_is_alt)
        c.var_hp1_is_alt = np.ascontiguousarray(h[keep].astype(np.uint8))
        c.var_ps = np.ascontiguousarray(np.asarray(ps)[keep].astype(np.int32))
        c.var_gt_kind = np.ones(len(keep), np.uint8)
        blob, off = b"", [0]
        for i in keep:
            o, e = int(self.var_str_off[i]), int(self.var_str_off[i + 1])
            blob += self.var_str[o:e]
            off.append(len(blob))
        c.var_str, c.var_str_off = blob, np.array(off, np.uint32)
        return c

    def somatic_union(self, seed=1, block=50, nor_frac=0.9, germ_in_tumor=0.15):
        """The contig as `somatic_haplotag` sees it: a union variant table (std::map<int, MultiGenomeVar>) built from the
        generator's truth.  NORMAL records = phased germline variants (true phase, artificial phase sets of `block`
        variants); TUMOR records = every somatic variant plus a share of the germline ones, with mixed genotypes.  Positions
        that end up in neither VCF are dropped.  Reads are shared with self."""
        import copy
        rng = np.random.default_rng(seed)
        n = self.n_var
        som = self.var_is_somatic != 0
        nor = (~som) & (rng.random(n) < nor_frac)
        tum = som | ((~som) & (rng.random(n) < germ_in_tumor))
        keep = np.nonzero(nor | tum)[0]
        c = copy.copy(self)
        c.n_var = len(keep)
        for k in ("var_pos", "var_ref0", "var_alt0", "var_ref_len", "var_alt_len", "var_hp1_is_alt", "var_is_somatic"):
            setattr(c, k, np.ascontiguousarray(getattr(self, k)[keep]))
        blk = (np.arange(c.n_var) // block) * block
        c.var_ps = np.ascontiguousarray((c.var_pos[blk] + 1).astype(np.int32))
        c.var_gt_kind = np.ones(c.n_var, np.uint8)
        blob, off = b"", [0]
        for i in keep:
            o, e = int(self.var_str_off[i]), int(self.var_str_off[i + 1])
            blob += self.var_str[o:e]
            off.append(len(blob))
        c.var_str, c.var_str_off = blob, np.array(off, np.uint32)
        c.nor_present = np.ascontiguousarray(nor[keep].astype(np.uint8))
        c.tum_present = np.ascontiguousarray(tum[keep].astype(np.uint8))
        m = c.n_var
        ksom = c.var_is_somatic != 0
        u = rng.random(m)
        c.tum_gt = np.where(ksom, np.where(u < 0.7, 2, np.where(u < 0.8, 3, 1)), 1 + (u * 3).astype(np.int64)).astype(np.uint8)
        c.tum_gt[c.tum_present == 0] = 0
        c.tum_hp1_is_alt = np.ascontiguousarray(c.var_hp1_is_alt.copy())
        c.tum_ps = np.where(c.tum_gt == 1, c.var_ps, -1).astype(np.int32)
        c.tum_ref0, c.tum_alt0 = c.var_ref0.copy(), c.var_alt0.copy()
        c.tum_ref_len, c.tum_alt_len = c.var_ref_len.copy(), c.var_alt_len.copy()
        # a few positions where the tumor VCF reports another ALT base than the normal VCF
        snp = (c.var_ref_len == 1) & (c.var_alt_len == 1) & (c.tum_present != 0) & (c.nor_present != 0) & (rng.random(m) < 0.1)
        tblob, toff = b"", [0]
        for i in range(m):
            o, e = int(c.var_str_off[i]), int(c.var_str_off[i + 1])
            rec = c.var_str[o:e]
            if snp[i]:
                alt = next(x for x in b"ACGT" if x != c.var_ref0[i] and x != c.var_alt0[i])
                c.tum_alt0[i] = alt
                rec = bytes([c.var_ref0[i], 0, alt, 0])
            tblob += rec
            toff.append(len(tblob))
        c.tum_str, c.tum_str_off = tblob, np.array(toff, np.uint32)
        # the caller's verdict (SomaticVarCaller::getSomaticFlag): most true somatic positions, a few germline tumor-only ones
        c.is_somatic = (((ksom & (rng.random(m) < 0.85)) | (~ksom & (c.nor_present == 0) & (rng.random(m) < 0.05)))
                        & (c.tum_present != 0)).astype(np.uint8)
        c.derive_hp = np.where(c.is_somatic != 0, rng.integers(0, 3, m), 0).astype(np.int8)
        return c

    def tumor_struct(self):
        P = _ffi.ptr
        return _ffi.LpsTumorVariants(n=self.n_var, nor_present=P(self.nor_present, _ffi.u8p), tum_present=P(self.tum_present, _ffi.u8p),
                                     ref0=P(self.tum_ref0, _ffi.u8p), alt0=P(self.tum_alt0, _ffi.u8p),
                                     ref_len=P(self.tum_ref_len, _ffi.u16p), alt_len=P(self.tum_alt_len, _ffi.u16p),
                                     gt_kind=P(self.tum_gt, _ffi.u8p), hp1_is_alt=P(self.tum_hp1_is_alt, _ffi.u8p),
                                     ps=P(self.tum_ps, _ffi.i32p), is_somatic=P(self.is_somatic, _ffi.u8p),
                                     derive_hp=P(self.derive_hp, _ffi.i8p))

    def with_reads_of(self, other):
        """Same variant tables, the read batch of `other` (a contig generated with the same seed and another read_seed)."""
        import copy
        c = copy.copy(self)
        for k in ("n_reads", "ref_start", "l_qseq", "n_cigar", "cigar_off", "seq_off", "qual_off", "flag", "mapq", "name_rank", "hap",
                  "cigar", "seq4", "qual", "names"):
            setattr(c, k, getattr(other, k))
        return c

    def batch_struct(self):
        P = _ffi.ptr
        return _ffi.LpsReadBatch(n_reads=self.n_reads, ref_start=P(self.ref_start, _ffi.i32p),
                                 l_qseq=P(self.l_qseq, _ffi.i32p), n_cigar=P(self.n_cigar, _ffi.u32p),
                                 cigar_off=P(self.cigar_off, _ffi.u64p), seq_off=P(self.seq_off, _ffi.u64p),
                                 qual_off=P(self.qual_off, _ffi.u64p), flag=P(self.flag, _ffi.u16p),
                                 mapq=P(self.mapq, _ffi.u8p), name_rank=P(self.name_rank, _ffi.i32p),
                                 cigar=P(self.cigar, _ffi.u32p), cigar_len=len(self.cigar),
                                 seq4=P(self.seq4, _ffi.u8p), seq_bytes=len(self.seq4), qual=P(self.qual, _ffi.u8p),
                                 qual_bytes=len(self.qual))

    def pack_cigar16(self):
        """The compact wire format of the CIGAR stream (include/lps.h, lps_read_batch.cigar16), made by the library's own
        lps_pack_cigar16 as the host loop would while appending records.  Returns (cigar16, long_len, long_at)."""
        n = len(self.cigar)
        cap = int(((self.cigar >> 4) >= 0xFFF).sum())
        c16 = np.zeros(n, np.uint16)
        long_len, long_at = np.zeros(max(cap, 1), np.uint32), np.zeros(max(cap, 1), np.uint64)
        n_long = C.c_uint64(0)
        P = _ffi.ptr
        rc = _ffi.load_library().lps_pack_cigar16(P(self.cigar, _ffi.u32p), n, 0, P(c16, _ffi.u16p), P(long_len, _ffi.u32p),
                                                  P(long_at, _ffi.u64p), cap, C.byref(n_long))
        if rc != 0 or n_long.value != cap:
            raise RuntimeError(f"lps_pack_cigar16 failed: rc {rc}, {n_long.value} long ops of {cap}")
        return c16, long_len[:cap], long_at[:cap]

    def batch_struct16(self):
        """batch_struct() with the CIGAR in the compact wire format; the arrays are kept alive on the object."""
        self._c16 = self.pack_cigar16()
        b = self.batch_struct()
        P = _ffi.ptr
        b.cigar = C.cast(None, _ffi.u32p)
        b.cigar16 = P(self._c16[0], _ffi.u16p)
        b.n_cigar_long = len(self._c16[1])
        if b.n_cigar_long:
            b.cigar_long_len = P(self._c16[1], _ffi.u32p)
            b.cigar_long_at = P(self._c16[2], _ffi.u64p)
        return b

    def pack_cigar8(self):
        """The 8-bit wire format of the CIGAR stream (include/lps.h, lps_read_batch.cigar8), made by the library's own lps_pack_cigar8
        as the host loop would while appending records.  Returns (cigar8, esc16, esc_blk, long_len, long_at)."""
        n = len(self.cigar)
        op, ln = self.cigar & 15, self.cigar >> 4
        short = ((op == 0) & (ln >= 1) & (ln <= 128)) | ((op == 1) & (ln >= 1) & (ln <= 56)) | ((op == 2) & (ln >= 1) & (ln <= 56))
        n_esc_want, n_long_want = int((~short).sum()), int(((~short) & (ln >= 0xFFF)).sum())
        c8 = np.zeros(max(n, 1), np.uint8)
        esc16 = np.zeros(max(n_esc_want, 1), np.uint16)
        esc_blk = np.zeros(n // 256 + 2, np.uint32)
        long_len, long_at = np.zeros(max(n_long_want, 1), np.uint32), np.zeros(max(n_long_want, 1), np.uint64)
        n_esc, n_long = C.c_uint64(0), C.c_uint64(0)
        P = _ffi.ptr
        rc = _ffi.load_library().lps_pack_cigar8(P(self.cigar, _ffi.u32p), n, 0, P(c8, _ffi.u8p), P(esc16, _ffi.u16p), n_esc_want, C.byref(n_esc),
                                                 P(esc_blk, _ffi.u32p), P(long_len, _ffi.u32p), P(long_at, _ffi.u64p), n_long_want, C.byref(n_long))
        if rc != 0 or n_esc.value != n_esc_want or n_long.value != n_long_want:
            raise RuntimeError(f"lps_pack_cigar8 failed: rc {rc}, {n_esc.value} of {n_esc_want} escapes, {n_long.value} of {n_long_want} long ops")
        return c8[:n], esc16[:n_esc_want], esc_blk, long_len[:n_long_want], long_at[:n_long_want]

    def batch_struct8(self):
        """batch_struct() with the CIGAR in the 8-bit wire format; the arrays are kept alive on the object."""
        self._c8 = self.pack_cigar8()
        c8, esc16, esc_blk, long_len, long_at = self._c8
        b = self.batch_struct()
        P = _ffi.ptr
        b.cigar = C.cast(None, _ffi.u32p)
        b.cigar8 = P(c8 if len(c8) else np.zeros(1, np.uint8), _ffi.u8p)
        b.cigar_esc_blk = P(esc_blk, _ffi.u32p)
        b.n_cigar_esc = len(esc16)
        if len(esc16):
            b.cigar_esc16 = P(esc16, _ffi.u16p)
        b.n_cigar_long = len(long_len)
        if len(long_len):
            b.cigar_long_len = P(long_len, _ffi.u32p)
            b.cigar_long_at = P(long_at, _ffi.u64p)
        return b

    def pack_sq(self, threads=8):
        """SEQ + QUAL as interleaved rows (include/lps.h, lps_read_batch.sq), made by the library's own lps_pack_sq_batch as the host
        loop would while appending records.  Returns (sq, sq_off): the stream and the 16-byte aligned row offset of every read."""
        lq = np.maximum(self.l_qseq.astype(np.int64), 0)
        rows = (lq + 9) // 10 * 16
        sq_off = np.zeros(self.n_reads, np.uint64)
        if self.n_reads:
            sq_off[1:] = np.cumsum(rows)[:-1].astype(np.uint64)
        sq = np.zeros(max(int(rows.sum()), 16), np.uint8)
        lib, P = _ffi.load_library(), _ffi.ptr
        l_qseq, seq_off, qual_off = (np.ascontiguousarray(a) for a in (self.l_qseq, self.seq_off, self.qual_off))

        def part(k):
            r0, r1 = self.n_reads * k // threads, self.n_reads * (k + 1) // threads
            if r1 > r0:
                rc = lib.lps_pack_sq_batch(r1 - r0, P(l_qseq[r0:r1], _ffi.i32p), P(seq_off[r0:r1], _ffi.u64p), P(qual_off[r0:r1], _ffi.u64p),
                                           P(self.seq4, _ffi.u8p), P(self.qual, _ffi.u8p), P(sq_off[r0:r1], _ffi.u64p), P(sq, _ffi.u8p))
                if rc != 0:
                    raise RuntimeError(f"lps_pack_sq_batch failed: rc {rc}")
        if threads > 1 and self.n_reads >= 64:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(part, range(threads)))
        else:
            threads = 1
            part(0)
        return sq, sq_off

    def batch_struct_sq(self):
        """batch_struct8() with SEQ + QUAL as interleaved rows (phase calls only); the arrays are kept alive on the object."""
        self._sq = self.pack_sq()
        b = self.batch_struct8()
        P = _ffi.ptr
        b.seq4, b.qual, b.qual_off = C.cast(None, _ffi.u8p), C.cast(None, _ffi.u8p), C.cast(None, _ffi.u64p)
        b.seq_bytes = b.qual_bytes = 0
        b.seq_off = P(self._sq[1], _ffi.u64p)
        b.sq, b.sq_bytes = P(self._sq[0], _ffi.u8p), len(self._sq[0])
        return b

    def name(self, i):
        s = self.names[i * self.NAME_STRIDE:(i + 1) * self.NAME_STRIDE]
        return s[:s.index(b"\0")].decode()

    def variant_strings(self, i):
        o = int(self.var_str_off[i])
        rl = int(self.var_ref_len[i])
        al = int(self.var_alt_len[i])
        return self.var_str[o:o + rl].decode(), self.var_str[o + rl + 1:o + rl + 1 + al].decode()
