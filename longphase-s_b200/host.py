"""Host-side mirror of the reference's phase interface over the C ABI (include/lps.h).

The names follow the reference seams this path replaces:
  BamParser.direct_detect_alleles  <- src/phase/ParsingBam.h:209
  VairiantGraph.addEdge / phasingProcess / exportResult  <- src/phase/PhasingGraph.h:187-192
Everything here is ctypes plumbing: the compute is in liblps_b200.so (CUDA, sm_100a) and there is no
CPU fallback — constructing a context without a usable GPU raises.
"""
import ctypes as C

import numpy as np

from . import _ffi


class LpsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lps error {code}: {msg}")
        self.code = code


class Context:
    """One context per (GPU, host thread) — lps_ctx_create / lps_ctx_destroy."""

    def __init__(self, device=0):
        self.lib = _ffi.load_library()
        h = C.c_void_p()
        rc = self.lib.lps_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise LpsError(rc, "lps_ctx_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self._keep = []

    def close(self):
        if self.h:
            self.lib.lps_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise LpsError(rc, self.lib.lps_last_error(self.h).decode())

    # ---- per-contig static data ----
    def set_reference(self, ref_bytes):
        self._check(self.lib.lps_contig_set_reference(self.h, ref_bytes, len(ref_bytes)))

    def set_variants(self, variants_struct, is_ont):
        self._check(self.lib.lps_contig_set_variants(self.h, C.byref(variants_struct), int(is_ont)))

    def notes(self):
        o = _ffi.LpsVariantNotes()
        self._check(self.lib.lps_contig_get_notes(self.h, C.byref(o)))
        g = _ffi.as_np
        return dict(homopolymer=g(o.homopolymer, o.n, np.uint8), is_danger=g(o.is_danger, o.n, np.uint8),
                    filtered=g(o.filtered, o.n, np.uint8))

    def submit(self, batch_struct):
        self._check(self.lib.lps_batch_submit(self.h, C.byref(batch_struct)))

    def submit_device(self, batch_struct):
        self._check(self.lib.lps_batch_submit_device(self.h, C.byref(batch_struct)))

    # ---- phase ----
    def call_alleles(self, params, want_host=True):
        o = _ffi.LpsCalls()
        self._check(self.lib.lps_phase_call_alleles(self.h, C.byref(params), int(want_host), C.byref(o)))
        g = _ffi.as_np
        res = dict(n_reads=o.n_reads, n_calls=int(o.n_calls), clip_pos=g(o.clip_pos, o.n_clips, np.int32),
                   clip_front=g(o.clip_front, o.n_clips, np.int32), clip_back=g(o.clip_back, o.n_clips, np.int32))
        if want_host:
            res.update(call_off=g(o.call_off, o.n_reads + 1, np.uint64), calls=g(o.calls, o.n_calls, _ffi.CALL_DTYPE),
                       read_status=g(o.read_status, o.n_reads, np.uint8))
        return res

    def build_edges(self, params, want_host=True):
        o = _ffi.LpsEdges()
        self._check(self.lib.lps_phase_build_edges(self.h, C.byref(params), int(want_host), C.byref(o)))
        g = _ffi.as_np
        res = dict(n_nodes=o.n_nodes, window=o.window, node_var=g(o.node_var, o.n_nodes, np.int32),
                   node_type=g(o.node_type, o.n_nodes, np.uint8), n_contrib=int(o.n_contrib),
                   n_contrib_far=int(o.n_contrib_far))
        if want_host:
            res["weights"] = g(o.weights, o.n_nodes * o.window * 4, np.float32).reshape(o.n_nodes, o.window, 4)
        return res

    def _result(self, o):
        g = _ffi.as_np
        return dict(ps=g(o.ps, o.n_variants, np.int32), hap_ref=g(o.hap_ref, o.n_variants, np.int8),
                    read_hp=g(o.read_hp, o.n_reads, np.int8),
                    hp_counts=g(o.hp_counts, o.n_variants * 4, np.int32).reshape(o.n_variants, 4),
                    ps_sweep=g(o.ps_sweep, o.n_variants, np.int32), hap_ref_sweep=g(o.hap_ref_sweep, o.n_variants, np.int8))

    def solve(self, params):
        o = _ffi.LpsPhaseResult()
        self._check(self.lib.lps_phase_solve(self.h, C.byref(params), C.byref(o)))
        return self._result(o)

    def phase_contig(self, params, copy=True):
        """lps_phase_contig.  copy=False skips the numpy copies of the result (the library has already brought it to host memory
        it owns; a caller that only wants the last of many steps saves the Python work, which holds the GIL)."""
        o = _ffi.LpsPhaseResult()
        self._check(self.lib.lps_phase_contig(self.h, C.byref(params), C.byref(o)))
        return self._result(o) if copy else None

    def event_record(self, slot):
        self._check(self.lib.lps_event_record(self.h, int(slot)))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._check(self.lib.lps_event_elapsed_ms(self.h, int(a), int(b), C.byref(ms)))
        return ms.value

    # ---- haplotag ----
    def tag_reads(self, tparams, want_calls=True):
        o = _ffi.LpsTagResult()
        self._check(self.lib.lps_tag_reads(self.h, C.byref(tparams), int(want_calls), C.byref(o)))
        g = _ffi.as_np
        n = o.n_reads
        res = dict(category=g(o.category, n, np.uint8), hp=g(o.hp, n, np.int8), ps=g(o.ps, n, np.int32), pq=g(o.pq, n, np.int32),
                   h1=g(o.h1, n, np.int32), h2=g(o.h2, n, np.int32), stats={k: getattr(o, k) for k in _ffi.TAG_COUNTERS})
        if want_calls:
            res.update(call_off=g(o.call_off, n + 1, np.uint64), calls=g(o.calls, o.n_calls, _ffi.CALL_DTYPE))
        return res

    # ---- somatic family ----
    def set_tumor_variants(self, tumor_struct):
        self._check(self.lib.lps_contig_set_tumor_variants(self.h, C.byref(tumor_struct)))

    @staticmethod
    def _read_tags(t):
        g = _ffi.as_np
        n = t.n_reads
        return dict(category=g(t.category, n, np.uint8), read_hp=g(t.read_hp, n, np.int8), ps=g(t.ps, n, np.int32), pq=g(t.pq, n, np.int32),
                    h1=g(t.h1, n, np.int32), h2=g(t.h2, n, np.int32), h3=g(t.h3, n, np.int32), n_ps=g(t.n_ps, n, np.uint8),
                    end_pos=g(t.end_pos, n, np.int32), read_len=g(t.read_len, n, np.int32))

    def _extract(self, fn, tparams, tumor):
        o = _ffi.LpsExtractResult()
        self._check(fn(self.h, C.byref(tparams), C.byref(o)))
        g = _ffi.as_np
        nt = o.n_tum
        res = self._read_tags(o.reads)
        res.update(n_tum=nt, tum_var=g(o.tum_var, nt, np.int32), pos_base=g(o.pos_base, nt * 15, np.int32).reshape(nt, 15),
                   read_hp_count=g(o.read_hp_count, nt * 9, np.int32).reshape(nt, 9),
                   ratios_f=g(o.ratios_f, nt * len(_ffi.RF_FIELDS), np.float32).reshape(nt, len(_ffi.RF_FIELDS)),
                   ratios_d=g(o.ratios_d, nt * len(_ffi.RD_FIELDS), np.float64).reshape(nt, len(_ffi.RD_FIELDS)),
                   case_read_count=g(o.case_read_count, nt, np.int32))
        if tumor:
            res.update(somatic_read_hp_count=g(o.somatic_read_hp_count, nt * 9, np.int32).reshape(nt, 9),
                       case_count=g(o.case_count, nt * 6, np.int32).reshape(nt, 6),
                       allele_count=g(o.allele_count, nt * 2, np.int32).reshape(nt, 2),
                       window_hist=g(o.window_hist, nt * 2 * _ffi.WINDOW_BINS, np.int32).reshape(nt, 2, _ffi.WINDOW_BINS),
                       n_window_items=int(o.n_window_items), call_off=g(o.call_off, o.reads.n_reads + 1, np.uint64),
                       calls=g(o.calls, o.n_calls, _ffi.CALL_DTYPE))
        return res

    def extract_normal(self, tparams):
        return self._extract(self.lib.lps_extract_normal, tparams, False)

    def extract_tumor(self, tparams):
        return self._extract(self.lib.lps_extract_tumor, tparams, True)

    def somatic_tag_reads(self, tparams, want_calls=True):
        o = _ffi.LpsSomaticTagResult()
        self._check(self.lib.lps_somatic_tag_reads(self.h, C.byref(tparams), int(want_calls), C.byref(o)))
        g = _ffi.as_np
        n, nt = o.reads.n_reads, o.n_tum
        res = self._read_tags(o.reads)
        res.update(hp_before=g(o.hp_before, n, np.int8), derive_similarity=g(o.derive_similarity, n, np.float32), n_tum=nt,
                   tum_var=g(o.tum_var, nt, np.int32), cover_start=g(o.cover_start, nt, np.int32), cover_end=g(o.cover_end, nt, np.int32),
                   stats={**{k: getattr(o, k) for k in _ffi.SOMATIC_COUNTERS}, **{f"hp{k}": o.total_hp[k] for k in range(9)}})
        for k in ("hp_before_count", "hp_after_count", "h3_before_count", "h3_after_count"):
            res[k] = g(getattr(o, k), nt * 9, np.int32).reshape(nt, 9)
        if want_calls:
            res.update(call_off=g(o.call_off, n + 1, np.uint64), calls=g(o.calls, o.n_calls, _ffi.CALL_DTYPE))
        return res

    # ---- BGZF (htslib bgzf_read_block / inflate_block) ----
    def bgzf_inflate(self, data, check_crc=True):
        """Inflates every BGZF block of `data` (bytes / uint8 array holding whole members) on the device; returns the bytes."""
        buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        blocks, out_bytes = bgzf_scan(buf)
        out = np.zeros(max(out_bytes, 1), np.uint8)
        self._check(self.lib.lps_bgzf_inflate(self.h, _ffi.ptr(buf, _ffi.u8p), len(buf), blocks.ctypes.data_as(C.POINTER(_ffi.LpsBgzfBlock)),
                                              len(blocks), _ffi.ptr(out, _ffi.u8p), out_bytes, int(check_crc)))
        return out[:out_bytes]

    # ---- BGZF (htslib bgzf_write / deflate_block) ----
    def bgzf_deflate(self, data, block_bytes=0xff00):
        """Deflates `data` on the device into BGZF members of block_bytes input bytes each (no EOF marker); returns the bytes."""
        buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        cap = int(self.lib.lps_bgzf_deflate_bound(len(buf), block_bytes))
        out = np.zeros(max(cap, 1), np.uint8)
        n = C.c_uint64(0)
        self._check(self.lib.lps_bgzf_deflate(self.h, _ffi.ptr(buf if len(buf) else np.zeros(1, np.uint8), _ffi.u8p), len(buf), block_bytes,
                                              _ffi.ptr(out, _ffi.u8p), cap, C.byref(n)))
        return out[:n.value]

    def stats(self):
        s = _ffi.LpsStats()
        self._check(self.lib.lps_get_stats(self.h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}


def bgzf_scan(buf):
    """lps_bgzf_scan: the block table (numpy, _ffi.BGZF_BLOCK_DTYPE) and the inflated size of a BGZF byte range.  Host only."""
    lib = _ffi.load_library()
    buf = np.ascontiguousarray(buf, np.uint8)
    n, total = C.c_uint64(0), C.c_uint64(0)
    rc = lib.lps_bgzf_scan(_ffi.ptr(buf, _ffi.u8p), len(buf), None, 0, C.byref(n), C.byref(total))
    if rc != 0:
        raise LpsError(rc, "lps_bgzf_scan: malformed BGZF data")
    blocks = np.zeros(max(n.value, 1), _ffi.BGZF_BLOCK_DTYPE)
    rc = lib.lps_bgzf_scan(_ffi.ptr(buf, _ffi.u8p), len(buf), blocks.ctypes.data_as(C.POINTER(_ffi.LpsBgzfBlock)), n.value, C.byref(n),
                           C.byref(total))
    if rc != 0:
        raise LpsError(rc, "lps_bgzf_scan failed")
    return blocks[:n.value], int(total.value)


class BamParser:
    """Mirror of reference BamParser (src/phase/ParsingBam.h:187-210) for one contig: the constructor takes
    the contig's variant table and reference string, direct_detect_alleles takes the decoded alignments."""

    def __init__(self, ctx, contig, params):
        self.ctx, self.params = ctx, params
        ctx.set_reference(contig.ref if params.have_reference else b"")
        self._vs = contig.variants_struct()
        ctx.set_variants(self._vs, params.is_ont)

    def direct_detect_alleles(self, contig, want_host=True):
        self._bs = contig.batch_struct()
        self.ctx.submit(self._bs)
        return self.ctx.call_alleles(self.params, want_host)


class VairiantGraph:
    """Mirror of reference VairiantGraph (src/phase/PhasingGraph.h:128-197); the misspelling is the reference's."""

    def __init__(self, ctx, params):
        self.ctx, self.params = ctx, params

    def addEdge(self, want_host=True):
        return self.ctx.build_edges(self.params, want_host)

    def phasingProcess(self):
        self.result = self.ctx.solve(self.params)
        return self.result

    def exportResult(self, contig):
        """-> list of (pos0, 'a|b', PS) like PhasingResult (src/phase/PhasingGraph.cpp:1049-1077)."""
        r = self.result
        out = []
        for i in np.nonzero(r["ps"])[0]:
            h = int(r["hap_ref"][i])
            out.append((int(contig.var_pos[i]), f"{h}|{1 - h}", int(r["ps"][i])))
        return out


class GermlineHaplotagChrProcessor:
    """Mirror of reference GermlineHaplotagChrProcessor (src/haplotag/HaplotagProcess.h:98-152): processSingleChrom's
    dispatch plus processRead (judgeHaplotype -> HP/PS/PQ) for every alignment of one contig."""

    def __init__(self, ctx, contig, tparams):
        self.ctx, self.tparams = ctx, tparams
        ctx.set_reference(contig.ref if tparams.have_reference else b"")
        self._vs = contig.variants_struct()
        ctx.set_variants(self._vs, 0)

    def processSingleChrom(self, contig, want_calls=True):
        self._bs = contig.batch_struct()
        self.ctx.submit(self._bs)
        return self.ctx.tag_reads(self.tparams, want_calls)


class _SomaticChrProcessor:
    """Shared set-up of the three somatic processors: reference string, union variant map (NORMAL + TUMOR records)."""

    def __init__(self, ctx, contig, tparams):
        self.ctx, self.tparams = ctx, tparams
        ctx.set_reference(contig.ref if tparams.have_reference else b"")
        self._vs, self._ts = contig.variants_struct(), contig.tumor_struct()
        ctx.set_variants(self._vs, 0)
        ctx.set_tumor_variants(self._ts)

    def _submit(self, contig):
        self._bs = contig.batch_struct()
        self.ctx.submit(self._bs)


class ExtractNorDataChrProcessor(_SomaticChrProcessor):
    """Mirror of reference ExtractNorDataChrProcessor (src/somatic_haplotag/SomaticVarCaller.h:235-259): processSingleChrom
    over the NORMAL BAM's alignments of one contig -> PosBase counters per tumor position."""

    def processSingleChrom(self, contig):
        self._submit(contig)
        return self.ctx.extract_normal(self.tparams)


class ExtractTumDataChrProcessor(_SomaticChrProcessor):
    """Mirror of reference ExtractTumDataChrProcessor (SomaticVarCaller.h:332-373) over the TUMOR BAM's alignments."""

    def processSingleChrom(self, contig):
        self._submit(contig)
        return self.ctx.extract_tumor(self.tparams)


class SomaticHaplotagChrProcessor(_SomaticChrProcessor):
    """Mirror of reference SomaticHaplotagChrProcessor (src/somatic_haplotag/SomaticHaplotagProcess.h:67-122): HP:Z / PS:i / PQ:i
    of every TUMOR alignment plus the per-position haplotype distributions before / after inheritance."""

    def processSingleChrom(self, contig, want_calls=True):
        self._submit(contig)
        return self.ctx.somatic_tag_reads(self.tparams, want_calls)


class TumorPurityEstimator:
    """Mirror of reference TumorPurityEstimator (src/somatic_haplotag/TumorPurityEstimator.h:276-349): built from the results of the NORMAL
    and TUMOR extract passes of every contig (chrVec order), estimateTumorPurity() returns the purity; `result` keeps the box-plot
    values, the valley threshold and the filter counts of the `_purity.out` log, `used` the per-position statisticPurity flags."""

    def __init__(self, normal_results, tumor_results):
        cat = lambda rs, k, j: np.ascontiguousarray(np.concatenate([r[k][:, j] for r in rs]))   # noqa: E731
        self.t_imb = cat(tumor_results, "ratios_d", 0).astype(np.float64)
        self.n_imb = cat(normal_results, "ratios_d", 0).astype(np.float64)
        self.n_pct = cat(normal_results, "ratios_d", 1).astype(np.float64)
        self.n_h1 = cat(normal_results, "read_hp_count", 1).astype(np.int32)
        self.n_h2 = cat(normal_results, "read_hp_count", 2).astype(np.int32)
        self.used = np.zeros(len(self.t_imb), np.uint8)

    def estimateTumorPurity(self):
        P = _ffi.ptr
        i = _ffi.LpsPurityInput(n=len(self.t_imb), tumor_germline_imbalance=P(self.t_imb, _ffi.f64p), normal_germline_imbalance=P(self.n_imb, _ffi.f64p),
                                normal_pct_germline_hp=P(self.n_pct, _ffi.f64p), normal_h1=P(self.n_h1, _ffi.i32p), normal_h2=P(self.n_h2, _ffi.i32p),
                                used=P(self.used, _ffi.u8p))
        o = _ffi.LpsPurityResult()
        rc = _ffi.load_library().lps_estimate_purity(C.byref(i), C.byref(o))
        if rc != 0:
            raise LpsError(rc, "lps_estimate_purity failed")
        self.result = {f: getattr(o, f) for f, _ in o._fields_}
        return o.purity


class SomaticVarCaller:
    """Mirror of the calling stage of reference SomaticVarCaller::variantCalling (src/somatic_haplotag/SomaticVarCaller.cpp:816-866)
    for one contig: built from the union contig (positions and types of the TUMOR records), variantCalling() takes the results
    of the NORMAL and TUMOR extract passes (dicts as ExtractNorDataChrProcessor / ExtractTumDataChrProcessor return them) and the
    tumor purity, and returns per tumor slot the filter flags, isSomaticVariant / somaticReadDeriveByHP (getSomaticFlag) and per
    alignment the calibrated read haplotype."""

    def __init__(self, contig, tparams, enableFilter=True):
        self.contig, self.tparams, self.enableFilter = contig, tparams, bool(enableFilter)

    @staticmethod
    def _as_struct(res, keep):
        P = _ffi.ptr
        def arr(k, dt):
            a = np.ascontiguousarray(res[k], dt) if k in res and res[k] is not None else None
            keep.append(a)
            return a
        n = len(res["h1"])
        reads = _ffi.LpsReadTags(n_reads=n, category=P(arr("category", np.uint8), _ffi.u8p), read_hp=P(arr("read_hp", np.int8), _ffi.i8p),
                                 h1=P(arr("h1", np.int32), _ffi.i32p), h2=P(arr("h2", np.int32), _ffi.i32p), h3=P(arr("h3", np.int32), _ffi.i32p),
                                 n_ps=P(arr("n_ps", np.uint8), _ffi.u8p), end_pos=P(arr("end_pos", np.int32), _ffi.i32p))
        o = _ffi.LpsExtractResult(n_tum=len(res["tum_var"]), tum_var=P(arr("tum_var", np.int32), _ffi.i32p),
                                  pos_base=P(arr("pos_base", np.int32), _ffi.i32p), read_hp_count=P(arr("read_hp_count", np.int32), _ffi.i32p),
                                  reads=reads, ratios_f=P(arr("ratios_f", np.float32), _ffi.f32p),
                                  case_read_count=P(arr("case_read_count", np.int32), _ffi.i32p))
        if "somatic_read_hp_count" in res and res["somatic_read_hp_count"] is not None and "calls" in res:
            calls = np.ascontiguousarray(res["calls"], _ffi.CALL_DTYPE)
            keep.append(calls)
            o.somatic_read_hp_count = P(arr("somatic_read_hp_count", np.int32), _ffi.i32p)
            o.window_hist = P(arr("window_hist", np.int32), _ffi.i32p)
            o.call_off = P(arr("call_off", np.uint64), _ffi.u64p)
            o.n_calls = len(calls)
            o.calls = calls.ctypes.data_as(C.POINTER(_ffi.LpsCall))
        return o

    def variantCalling(self, normal_result, tumor_result, tumorPurity):
        c = self.contig
        tum_var = np.ascontiguousarray(tumor_result["tum_var"], np.int32)
        nt, nr = len(tum_var), len(tumor_result["h1"])
        pos = np.ascontiguousarray(c.var_pos[tum_var], np.int32)
        rl, al = c.tum_ref_len[tum_var], c.tum_alt_len[tum_var]
        callable_ = np.ascontiguousarray(((rl == 1) | (al == 1)).astype(np.uint8))          # SNP, insertion or deletion (VarData::setVariantType)
        keep = []
        sn, st = self._as_struct(normal_result, keep), self._as_struct(tumor_result, keep)
        P = _ffi.ptr
        i = _ffi.LpsSomaticCallInput(n_tum=nt, pos=P(pos, _ffi.i32p), callable=P(callable_, _ffi.u8p), normal=C.pointer(sn), tumor=C.pointer(st),
                                     purity=float(tumorPurity), enable_filter=int(self.enableFilter),
                                     percentage_threshold=self.tparams.percentage_threshold)
        r = dict(touched=np.zeros(nt, np.uint8), is_somatic=np.zeros(nt, np.uint8), derive_hp=np.zeros(nt, np.int8),
                 is_filter_out=np.zeros(nt, np.uint8), filtered_by=np.zeros((nt, len(_ffi.FILTER_FIELDS)), np.uint8),
                 in_dense_interval=np.zeros(nt, np.uint8), mean_alt_per_var_read=np.zeros(nt, np.float32), z_score=np.zeros(nt, np.float32),
                 interval_snp_count=np.zeros(nt, np.int32), min_distance=np.zeros(nt, np.int32), dense_alt_same_count=np.zeros(nt, np.int32),
                 read_hp=np.zeros(nr, np.int8), read_h3=np.zeros(nr, np.int32))
        types = dict(touched=_ffi.u8p, is_somatic=_ffi.u8p, derive_hp=_ffi.i8p, is_filter_out=_ffi.u8p, filtered_by=_ffi.u8p, in_dense_interval=_ffi.u8p,
                     mean_alt_per_var_read=_ffi.f32p, z_score=_ffi.f32p, interval_snp_count=_ffi.i32p, min_distance=_ffi.i32p,
                     dense_alt_same_count=_ffi.i32p, read_hp=_ffi.i8p, read_h3=_ffi.i32p)
        o = _ffi.LpsSomaticCallResult(**{k: P(v, types[k]) for k, v in r.items()})
        rc = _ffi.load_library().lps_somatic_call(C.byref(i), C.byref(o))
        if rc != 0:
            raise LpsError(rc, "lps_somatic_call: a tumor position without any read record (the reference prints [ERROR] and exits)"
                           if rc == -6 else "lps_somatic_call failed")
        r.update(tier=o.tier, n_somatic=o.n_somatic, tum_var=tum_var)
        return r
