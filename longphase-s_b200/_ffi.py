"""ctypes mirror of include/lps.h (the C ABI).  Plain pointers and sizes only."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblps_b200.so")

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
u16p = C.POINTER(C.c_uint16)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)


class LpsVariants(C.Structure):
    _fields_ = [("n", C.c_int32), ("pos", i32p), ("ref0", u8p), ("alt0", u8p), ("ref_len", u16p), ("alt_len", u16p),
                ("hp1_is_alt", u8p), ("ps", i32p), ("gt_kind", u8p)]


class LpsReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("ref_start", i32p), ("l_qseq", i32p), ("n_cigar", u32p),
                ("cigar_off", u64p), ("seq_off", u64p), ("qual_off", u64p), ("flag", u16p), ("mapq", u8p),
                ("name_rank", i32p), ("cigar", u32p), ("cigar_len", C.c_uint64), ("seq4", u8p),
                ("seq_bytes", C.c_uint64), ("qual", u8p), ("qual_bytes", C.c_uint64),
                ("cigar16", u16p), ("cigar_long_len", u32p), ("cigar_long_at", u64p), ("n_cigar_long", C.c_uint64),
                ("cigar8", u8p), ("cigar_esc16", u16p), ("n_cigar_esc", C.c_uint64), ("cigar_esc_blk", u32p),
                ("sq", u8p), ("sq_bytes", C.c_uint64)]


class LpsBgzfBlock(C.Structure):
    _fields_ = [("comp_off", C.c_uint64), ("comp_len", C.c_uint32), ("out_len", C.c_uint32), ("out_off", C.c_uint64),
                ("crc32", C.c_uint32), ("reserved_", C.c_uint32)]


BGZF_BLOCK_DTYPE = np.dtype([("comp_off", "<u8"), ("comp_len", "<u4"), ("out_len", "<u4"), ("out_off", "<u8"), ("crc32", "<u4"),
                             ("reserved_", "<u4")])


class LpsCall(C.Structure):
    _fields_ = [("var", C.c_int32), ("quality", C.c_int16), ("allele", C.c_int8), ("origin", C.c_int8)]


CALL_DTYPE = np.dtype([("var", "<i4"), ("quality", "<i2"), ("allele", "i1"), ("origin", "i1")])


class LpsPhaseParams(C.Structure):
    _fields_ = [("mapping_quality", C.c_int32), ("is_ont", C.c_int32), ("have_reference", C.c_int32),
                ("connect_adjacent", C.c_int32), ("base_quality", C.c_int32), ("distance", C.c_int32),
                ("edge_weight", C.c_double), ("edge_threshold", C.c_double), ("overlap_threshold", C.c_double),
                ("read_confidence", C.c_double), ("snp_confidence", C.c_double)]


def default_phase_params(is_ont=True):
    """Defaults of `longphase-s phase` (reference src/phase/Phasing.cpp:88-116)."""
    return LpsPhaseParams(mapping_quality=1, is_ont=int(is_ont), have_reference=1, connect_adjacent=35,
                          base_quality=12, distance=300000, edge_weight=0.1, edge_threshold=0.7,
                          overlap_threshold=0.2, read_confidence=0.65, snp_confidence=0.75)


class LpsCalls(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_calls", C.c_uint64), ("call_off", u64p), ("calls", C.POINTER(LpsCall)),
                ("read_status", u8p), ("n_clips", C.c_int32), ("clip_pos", i32p), ("clip_front", i32p),
                ("clip_back", i32p)]


class LpsVariantNotes(C.Structure):
    _fields_ = [("n", C.c_int32), ("homopolymer", u8p), ("is_danger", u8p), ("filtered", u8p)]


class LpsEdges(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_var", i32p), ("node_type", u8p), ("window", C.c_int32),
                ("weights", f32p), ("n_contrib", C.c_uint64), ("n_contrib_far", C.c_uint64)]


class LpsPhaseResult(C.Structure):
    _fields_ = [("n_variants", C.c_int32), ("ps", i32p), ("hap_ref", i8p), ("n_reads", C.c_int32),
                ("read_hp", i8p), ("hp_counts", i32p), ("ps_sweep", i32p), ("hap_ref_sweep", i8p)]


class LpsTagParams(C.Structure):
    _fields_ = [("mapping_quality", C.c_int32), ("mapq_filter", C.c_int32), ("tag_supplementary", C.c_int32),
                ("have_reference", C.c_int32), ("percentage_threshold", C.c_double)]


def default_tag_params():
    """Defaults of `longphase-s haplotag` (reference src/haplotag/Haplotag.cpp:60-72)."""
    return LpsTagParams(mapping_quality=1, mapq_filter=1, tag_supplementary=0, have_reference=1, percentage_threshold=0.6)


TAG_COUNTERS = ["total_alignment", "total_supplementary", "total_secondary", "total_unmapped", "total_tag", "total_untag",
                "total_lower_quality", "total_other_case", "total_empty_variant", "total_high_similarity",
                "total_without_variant", "total_hp1", "total_hp2", "total_hp0"]


class LpsTagResult(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("category", u8p), ("hp", i8p), ("ps", i32p), ("pq", i32p), ("h1", i32p), ("h2", i32p),
                ("n_calls", C.c_uint64), ("call_off", u64p), ("calls", C.POINTER(LpsCall))] + [(k, C.c_int64) for k in TAG_COUNTERS]


class LpsTumorVariants(C.Structure):
    _fields_ = [("n", C.c_int32), ("nor_present", u8p), ("tum_present", u8p), ("ref0", u8p), ("alt0", u8p), ("ref_len", u16p),
                ("alt_len", u16p), ("gt_kind", u8p), ("hp1_is_alt", u8p), ("ps", i32p), ("is_somatic", u8p), ("derive_hp", i8p)]


PB_FIELDS = ["alt", "A", "C", "G", "T", "unknown", "depth", "del", "mpq_alt", "mpq_A", "mpq_C", "mpq_G", "mpq_T", "mpq_unknown",
             "mpq_depth"]


class LpsReadTags(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("category", u8p), ("read_hp", i8p), ("ps", i32p), ("pq", i32p), ("h1", i32p), ("h2", i32p),
                ("h3", i32p), ("n_ps", u8p), ("end_pos", i32p), ("read_len", i32p)]


CASE_FIELDS = ["clean_hp3", "pure_h1_1", "pure_h2_1", "pure_h3", "mixed", "untag"]
WINDOW_BINS = 201


class LpsExtractResult(C.Structure):
    _fields_ = [("n_tum", C.c_int32), ("tum_var", i32p), ("pos_base", i32p), ("read_hp_count", i32p), ("reads", LpsReadTags),
                ("somatic_read_hp_count", i32p), ("case_count", i32p), ("allele_count", i32p), ("window_hist", i32p),
                ("n_window_items", C.c_uint64), ("ratios_f", f32p), ("ratios_d", C.POINTER(C.c_double)), ("case_read_count", i32p),
                ("n_calls", C.c_uint64), ("call_off", u64p), ("calls", C.POINTER(LpsCall))]


class LpsSomaticCallInput(C.Structure):
    _fields_ = [("n_tum", C.c_int32), ("pos", i32p), ("callable", u8p), ("normal", C.POINTER(LpsExtractResult)),
                ("tumor", C.POINTER(LpsExtractResult)), ("purity", C.c_double), ("enable_filter", C.c_int32),
                ("percentage_threshold", C.c_double)]


FILTER_FIELDS = ["tinc", "messy_read", "read_count", "hap_consistency", "variant_cluster", "dense_alt"]


class LpsSomaticCallResult(C.Structure):
    _fields_ = [("touched", u8p), ("is_somatic", u8p), ("derive_hp", i8p), ("is_filter_out", u8p), ("filtered_by", u8p),
                ("in_dense_interval", u8p), ("mean_alt_per_var_read", f32p), ("z_score", f32p), ("interval_snp_count", i32p),
                ("min_distance", i32p), ("dense_alt_same_count", i32p), ("read_hp", i8p), ("read_h3", i32p), ("tier", C.c_int32),
                ("n_somatic", C.c_int32)]


RF_FIELDS = ["vaf", "non_del_vaf", "mpq_vaf", "low_mpq_ratio", "del_ratio", "mixed_ratio", "pure_h1_1_ratio", "pure_h2_1_ratio", "pure_h3_ratio"]
RD_FIELDS = ["germline_imbalance", "pct_germline_hp", "allelic_imbalance", "somatic_imbalance"]
SOMATIC_COUNTERS = ["total_alignment", "total_supplementary", "total_secondary", "total_unmapped", "total_tag", "total_untag",
                    "total_lower_quality", "total_other_case", "total_empty_variant", "total_high_similarity", "total_cross_two_block",
                    "total_without_variant", "total_read_only_h3"]


class LpsSomaticTagResult(C.Structure):
    _fields_ = [("reads", LpsReadTags), ("hp_before", i8p), ("derive_similarity", f32p), ("n_tum", C.c_int32), ("tum_var", i32p),
                ("hp_before_count", i32p), ("hp_after_count", i32p), ("h3_before_count", i32p), ("h3_after_count", i32p),
                ("cover_start", i32p), ("cover_end", i32p), ("n_calls", C.c_uint64), ("call_off", u64p),
                ("calls", C.POINTER(LpsCall))] + [(k, C.c_int64) for k in SOMATIC_COUNTERS] + [("total_hp", C.c_int64 * 9)]


f64p = C.POINTER(C.c_double)


class LpsPurityInput(C.Structure):
    _fields_ = [("n", C.c_int32), ("tumor_germline_imbalance", f64p), ("normal_germline_imbalance", f64p), ("normal_pct_germline_hp", f64p),
                ("normal_h1", i32p), ("normal_h2", i32p), ("used", u8p)]


class LpsPurityResult(C.Structure):
    _fields_ = [("purity", C.c_double), ("ok", C.c_int32), ("read_count_threshold", C.c_int32), ("median", C.c_double), ("q1", C.c_double),
                ("q3", C.c_double), ("iqr", C.c_double), ("lower_whisker", C.c_double), ("upper_whisker", C.c_double),
                ("n_after_lcvf", C.c_int32), ("n_used", C.c_int32), ("filtered_normal_imbalance_zero", C.c_int32),
                ("filtered_tumor_imbalance_zero", C.c_int32), ("filtered_normal_imbalance_high", C.c_int32),
                ("filtered_normal_read_count", C.c_int32), ("filtered_pct_germline_hp", C.c_int32), ("filtered_valley", C.c_int32),
                ("filtered_outliers", C.c_int32), ("n_outliers_left", C.c_int32)]


class LpsStats(C.Structure):
    _fields_ = [("ms_call_alleles", C.c_float), ("ms_tag_reads", C.c_float), ("ms_build_edges", C.c_float), ("ms_read_correction", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("ms_kernel_call_alleles", C.c_float),
                ("ms_kernel_fold_edges", C.c_float), ("ms_kernel_window_diff", C.c_float), ("ms_wall_call_alleles", C.c_float),
                ("ms_wall_build_edges", C.c_float), ("ms_wall_solve", C.c_float), ("ms_host_filters", C.c_float),
                ("ms_host_sweep", C.c_float), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("sweep_simd", C.c_int32), ("ms_kernel_bgzf", C.c_float),
                ("ms_sweep", C.c_float), ("sweep_fallbacks", C.c_uint32), ("slow_path_contigs", C.c_uint32)]


# every symbol include/lps.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "lps_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "lps_ctx_destroy": (None, [C.c_void_p]),
    "lps_last_error": (C.c_char_p, [C.c_void_p]),
    "lps_version": (C.c_char_p, []),
    "lps_contig_set_reference": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "lps_contig_set_variants": (C.c_int, [C.c_void_p, C.POINTER(LpsVariants), C.c_int]),
    "lps_contig_get_notes": (C.c_int, [C.c_void_p, C.POINTER(LpsVariantNotes)]),
    "lps_batch_submit": (C.c_int, [C.c_void_p, C.POINTER(LpsReadBatch)]),
    "lps_batch_submit_device": (C.c_int, [C.c_void_p, C.POINTER(LpsReadBatch)]),
    "lps_pack_cigar16": (C.c_int, [u32p, C.c_uint64, C.c_uint64, u16p, u32p, u64p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "lps_pack_cigar8": (C.c_int, [u32p, C.c_uint64, C.c_uint64, u8p, u16p, C.c_uint64, C.POINTER(C.c_uint64), u32p, u32p, u64p, C.c_uint64,
                                  C.POINTER(C.c_uint64)]),
    "lps_sq_row_bytes": (C.c_uint64, [C.c_int32]),
    "lps_pack_sq": (C.c_int, [u8p, u8p, C.c_int32, u8p]),
    "lps_pack_sq_batch": (C.c_int, [C.c_int32, i32p, u64p, u64p, u8p, u8p, u64p, u8p]),
    "lps_sq_peek": (C.c_int, [u8p, C.c_int32, C.c_int32, u8p, u8p]),
    "lps_phase_call_alleles": (C.c_int, [C.c_void_p, C.POINTER(LpsPhaseParams), C.c_int, C.POINTER(LpsCalls)]),
    "lps_phase_build_edges": (C.c_int, [C.c_void_p, C.POINTER(LpsPhaseParams), C.c_int, C.POINTER(LpsEdges)]),
    "lps_phase_solve": (C.c_int, [C.c_void_p, C.POINTER(LpsPhaseParams), C.POINTER(LpsPhaseResult)]),
    "lps_phase_contig": (C.c_int, [C.c_void_p, C.POINTER(LpsPhaseParams), C.POINTER(LpsPhaseResult)]),
    "lps_sweep_votes": (C.c_int, [C.POINTER(LpsPhaseParams), C.c_int32, C.c_int32, i32p, u8p, u8p, i32p, i8p]),
    "lps_tag_reads": (C.c_int, [C.c_void_p, C.POINTER(LpsTagParams), C.c_int, C.POINTER(LpsTagResult)]),
    "lps_contig_set_tumor_variants": (C.c_int, [C.c_void_p, C.POINTER(LpsTumorVariants)]),
    "lps_extract_normal": (C.c_int, [C.c_void_p, C.POINTER(LpsTagParams), C.POINTER(LpsExtractResult)]),
    "lps_extract_tumor": (C.c_int, [C.c_void_p, C.POINTER(LpsTagParams), C.POINTER(LpsExtractResult)]),
    "lps_somatic_tag_reads": (C.c_int, [C.c_void_p, C.POINTER(LpsTagParams), C.c_int, C.POINTER(LpsSomaticTagResult)]),
    "lps_bgzf_scan": (C.c_int, [u8p, C.c_uint64, C.POINTER(LpsBgzfBlock), C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "lps_bgzf_inflate": (C.c_int, [C.c_void_p, u8p, C.c_uint64, C.POINTER(LpsBgzfBlock), C.c_uint64, u8p, C.c_uint64, C.c_int]),
    "lps_bgzf_inflate_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "lps_bgzf_deflate_bound": (C.c_uint64, [C.c_uint64, C.c_uint32]),
    "lps_bgzf_deflate": (C.c_int, [C.c_void_p, u8p, C.c_uint64, C.c_uint32, u8p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "lps_bgzf_deflate_block_host": (C.c_int, [u8p, C.c_uint32, u8p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "lps_set_blocking_sync": (C.c_int, [C.c_int, C.c_int]),
    "lps_estimate_purity": (C.c_int, [C.POINTER(LpsPurityInput), C.POINTER(LpsPurityResult)]),
    "lps_somatic_call": (C.c_int, [C.POINTER(LpsSomaticCallInput), C.POINTER(LpsSomaticCallResult)]),
    "lps_somatic_filter_params_of": (C.c_int, [C.c_double, C.c_void_p]),
    "lps_get_stats": (C.c_int, [C.c_void_p, C.POINTER(LpsStats)]),
    "lps_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "lps_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


def load_library():
    """Load liblps_b200.so.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no CPU fallback for the hot path")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(a, typ):
    """numpy array -> typed pointer (None -> NULL)."""
    if a is None:
        return C.cast(None, typ)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ)


def as_np(p, n, dtype):
    """typed pointer -> numpy copy of n elements."""
    n = int(n)
    if n == 0 or not p:
        return np.zeros(0, dtype=dtype)
    buf = C.cast(p, C.POINTER(C.c_uint8 * (n * np.dtype(dtype).itemsize))).contents
    return np.frombuffer(buf, dtype=dtype, count=n).copy()
