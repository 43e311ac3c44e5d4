"""ctypes wrappers of the oracle (liboracle.so) and of the reference tap (oracle/_ref/libref_tap.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import __graft_entry__ as _entry  # noqa: E402

_pkg = _entry.load_package()
_ffi = _pkg._ffi

ORACLE_LIB = os.path.join(HERE, "liboracle.so")
TAP_LIB = os.path.join(HERE, "_ref", "libref_tap.so")
REF_BIN = os.path.join(HERE, "_ref", "longphase-s")

i32p, u8p, i8p, u64p, f32p = _ffi.i32p, _ffi.u8p, _ffi.i8p, _ffi.u64p, _ffi.f32p
callp = C.POINTER(_ffi.LpsCall)


class OrcCalls(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_calls", C.c_uint64), ("call_off", u64p), ("calls", callp),
                ("read_status", u8p), ("n_clips", C.c_int32), ("clip_pos", i32p), ("clip_front", i32p),
                ("clip_back", i32p)]


class OrcGraph(C.Structure):
    _fields_ = [("n_aln", C.c_int32), ("aln_read", i32p), ("aln_off", u64p), ("aln_calls", callp),
                ("n_cnv", C.c_int32), ("cnv_start", i32p), ("cnv_end", i32p), ("n_nodes", C.c_int32),
                ("node_var", i32p), ("node_type", u8p), ("window", C.c_int32), ("weights", f32p),
                ("n_contrib", C.c_uint64), ("n_contrib_far", C.c_uint64), ("n_far_cells", C.c_uint64)]


class OrcSolution(C.Structure):
    _fields_ = [("n_variants", C.c_int32), ("ps_sweep", i32p), ("hap_ref_sweep", i8p), ("ps", i32p),
                ("hap_ref", i8p), ("n_aln", C.c_int32), ("read_hp", i8p), ("hp_counts", i32p)]


class TapIn(C.Structure):
    _fields_ = [("chr", C.c_char_p), ("ref", C.c_char_p), ("ref_len", C.c_int64), ("n_var", C.c_int32),
                ("var_pos", i32p), ("var_str_off", _ffi.u32p), ("var_str", C.c_char_p),
                ("batch", _ffi.LpsReadBatch), ("names", C.c_char_p), ("name_stride", C.c_int32),
                ("stop_after_calls", C.c_int32), ("p", _ffi.LpsPhaseParams)]


class TapCalls(C.Structure):
    _fields_ = [("n_aln", C.c_int32), ("read_idx", i32p), ("off", u64p), ("pos", i32p), ("allele", i32p),
                ("quality", i32p)]


class TapNodes(C.Structure):
    _fields_ = [("n", C.c_int32), ("pos", i32p), ("type", i32p), ("ps", i32p), ("hap_ref", i32p),
                ("hap_alt", i32p)]


class TapOut(C.Structure):
    _fields_ = [("stage_a", TapCalls), ("stage_b", TapCalls), ("stage_c", TapCalls), ("n_clips", C.c_int32),
                ("clip_pos", i32p), ("clip_front", i32p), ("clip_back", i32p), ("n_cnv", C.c_int32),
                ("cnv_start", i32p), ("cnv_end", i32p), ("n_empty_after_filter", C.c_int32),
                ("n_edge_nodes", C.c_int32), ("n_cells", C.c_int64), ("n_contrib", C.c_uint64),
                ("cell_a", i32p), ("cell_b", i32p), ("cell_which", u8p), ("cell_val", f32p),
                ("nodes_sweep", TapNodes), ("nodes_final", TapNodes), ("read_hp", i32p),
                ("n_result", C.c_int32), ("res_pos", i32p), ("res_block", i32p), ("res_hap_ref", i32p),
                ("res_hap_alt", i32p), ("t_get_snp", C.c_double), ("t_filter_snp", C.c_double),
                ("t_clip", C.c_double), ("t_add_edge", C.c_double), ("t_sweep", C.c_double),
                ("t_read_correction", C.c_double), ("t_export", C.c_double), ("n_result_calls", C.c_int64)]


class OrcTags(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("category", u8p), ("hp", i8p), ("ps", i32p), ("pq", i32p), ("h1", i32p), ("h2", i32p),
                ("n_calls", C.c_uint64), ("call_off", u64p), ("calls", callp)]


class TapTagIn(C.Structure):
    _fields_ = [("chr", C.c_char_p), ("ref", C.c_char_p), ("ref_len", C.c_int64), ("n_var", C.c_int32),
                ("var_pos", i32p), ("var_str_off", _ffi.u32p), ("var_str", C.c_char_p), ("var_hp1_is_alt", u8p),
                ("var_ps", i32p), ("batch", _ffi.LpsReadBatch), ("names", C.c_char_p), ("name_stride", C.c_int32),
                ("p", _ffi.LpsTagParams)]


class TapTagOut(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("category", u8p), ("hp", i32p), ("ps", i32p), ("pq", i32p), ("h1", i32p), ("h2", i32p),
                ("n_ps", i32p), ("var_off", u64p), ("var_pos", i32p), ("var_hp", i32p), ("ps_off", u64p), ("ps_id", i32p),
                ("ps_count", i32p), ("stats", C.c_int64 * 14), ("t_total", C.c_double)]


class OrcSomaticOut(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_tum", C.c_int32), ("tum_var", i32p), ("category", u8p), ("read_hp", i8p), ("hp_before", i8p),
                ("ps", i32p), ("pq", i32p), ("h1", i32p), ("h2", i32p), ("h3", i32p), ("n_ps", u8p), ("end_pos", i32p),
                ("read_len", i32p), ("derive_similarity", f32p), ("pos_base", i32p), ("read_hp_count", i32p),
                ("somatic_read_hp_count", i32p), ("case_count", i32p), ("allele_count", i32p), ("window_hist", i32p),
                ("hp_before_count", i32p), ("hp_after_count", i32p), ("h3_before_count", i32p), ("h3_after_count", i32p),
                ("cover_start", i32p), ("cover_end", i32p), ("ratios_f", f32p), ("ratios_d", C.POINTER(C.c_double)), ("case_read_count", i32p),
                ("n_window_items", C.c_uint64), ("n_calls", C.c_uint64),
                ("call_off", u64p), ("calls", callp)]


class TapSomIn(C.Structure):
    _fields_ = [("chr", C.c_char_p), ("ref", C.c_char_p), ("ref_len", C.c_int64), ("n_var", C.c_int32), ("var_pos", i32p),
                ("var_str_off", _ffi.u32p), ("var_str", C.c_char_p), ("var_hp1_is_alt", u8p), ("var_ps", i32p), ("nor_gt", u8p),
                ("nor_present", u8p), ("tum_present", u8p), ("tum_str_off", _ffi.u32p), ("tum_str", C.c_char_p), ("tum_gt", u8p),
                ("tum_hp1_is_alt", u8p), ("tum_ps", i32p), ("is_somatic", u8p), ("derive_hp", i8p), ("batch", _ffi.LpsReadBatch),
                ("names", C.c_char_p), ("name_stride", C.c_int32), ("p", _ffi.LpsTagParams), ("stats_out", C.POINTER(C.c_int64)),
                ("keep", C.c_void_p)]


class TapPurityOut(C.Structure):
    _fields_ = [("purity", C.c_double), ("threshold", C.c_int32), ("n_after_lcvf", C.c_int32), ("n_used", C.c_int32), ("median", C.c_double),
                ("q1", C.c_double), ("q3", C.c_double), ("iqr", C.c_double), ("lower_whisker", C.c_double), ("upper_whisker", C.c_double)]


class TapCallOut(C.Structure):
    _fields_ = [("n_tum", C.c_int32), ("n_reads", C.c_int32), ("touched", u8p), ("mean_alt", _ffi.f32p), ("z_score", _ffi.f32p),
                ("interval_snp_count", i32p), ("min_distance", i32p), ("dense_alt_same", i32p), ("in_dense", u8p), ("filtered_by", u8p),
                ("is_filter_out", u8p), ("high_con", u8p), ("derive_hp", i32p), ("is_somatic", u8p), ("flag_derive_hp", i32p),
                ("read_hp", i8p), ("read_h3", i32p), ("tier", C.c_int32)]


SOM_MODES = {"extract_normal": 0, "extract_tumor": 1, "somatic_tag": 2}

_orc = None
_tap = None


def oracle_lib():
    global _orc
    if _orc is None:
        lib = C.CDLL(ORACLE_LIB)
        lib.orc_annotate.argtypes = [C.c_char_p, C.c_int64, C.POINTER(_ffi.LpsVariants), C.c_int, u8p, u8p, u8p]
        lib.orc_call_alleles.argtypes = [C.POINTER(_ffi.LpsReadBatch), C.POINTER(_ffi.LpsVariants), u8p, u8p, u8p,
                                         C.c_int, C.POINTER(_ffi.LpsPhaseParams), C.POINTER(OrcCalls)]
        lib.orc_calls_free.argtypes = [C.POINTER(OrcCalls)]
        lib.orc_build_graph.argtypes = [C.POINTER(_ffi.LpsReadBatch), C.POINTER(_ffi.LpsVariants), u8p,
                                        C.POINTER(OrcCalls), C.POINTER(_ffi.LpsPhaseParams), C.POINTER(OrcGraph)]
        lib.orc_graph_free.argtypes = [C.POINTER(OrcGraph)]
        lib.orc_solve.argtypes = [C.POINTER(_ffi.LpsVariants), C.POINTER(OrcGraph), C.POINTER(_ffi.LpsPhaseParams),
                                  C.POINTER(OrcSolution)]
        lib.orc_solution_free.argtypes = [C.POINTER(OrcSolution)]
        lib.orc_tag_reads.argtypes = [C.POINTER(_ffi.LpsReadBatch), C.POINTER(_ffi.LpsVariants), u8p, C.POINTER(_ffi.LpsTagParams),
                                      C.POINTER(OrcTags)]
        lib.orc_tags_free.argtypes = [C.POINTER(OrcTags)]
        lib.orc_somatic.argtypes = [C.c_int, C.POINTER(_ffi.LpsReadBatch), C.POINTER(_ffi.LpsVariants), C.POINTER(_ffi.LpsTumorVariants),
                                    u8p, C.c_char_p, C.c_int64, C.POINTER(_ffi.LpsTagParams), C.POINTER(OrcSomaticOut)]
        lib.orc_somatic_free.argtypes = [C.POINTER(OrcSomaticOut)]
        _orc = lib
    return _orc


def tap_available():
    return os.path.exists(TAP_LIB)


def tap_lib():
    global _tap
    if _tap is None:
        lib = C.CDLL(TAP_LIB)
        lib.ref_tap_phase.argtypes = [C.POINTER(TapIn), C.POINTER(TapOut)]
        lib.ref_tap_phase.restype = C.c_int
        lib.ref_tap_phase_free.argtypes = [C.POINTER(TapOut)]
        lib.ref_tap_homopolymer.argtypes = [C.c_char_p, C.c_int64, C.c_int]
        lib.ref_tap_tag.argtypes = [C.POINTER(TapTagIn), C.POINTER(TapTagOut)]
        lib.ref_tap_tag.restype = C.c_int
        lib.ref_tap_tag_free.argtypes = [C.POINTER(TapTagOut)]
        lib.ref_tap_somatic.argtypes = [C.c_int, C.POINTER(TapSomIn), C.POINTER(OrcSomaticOut)]
        lib.ref_tap_somatic.restype = C.c_int
        lib.ref_tap_somatic_free.argtypes = [C.POINTER(OrcSomaticOut)]
        lib.ref_tap_state_new.restype = C.c_void_p
        lib.ref_tap_state_free.argtypes = [C.c_void_p]
        lib.ref_tap_purity.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(TapPurityOut)]
        lib.ref_tap_somatic_call.argtypes = [C.c_void_p, C.POINTER(TapSomIn), C.c_double, C.c_int, C.POINTER(TapCallOut)]
        lib.ref_tap_somatic_call.restype = C.c_int
        lib.ref_tap_somatic_call_free.argtypes = [C.POINTER(TapCallOut)]
        _tap = lib
    return _tap


g = _ffi.as_np


class Notes:
    def __init__(self, contig, is_ont):
        n = contig.n_var
        self.hom = np.zeros(n, np.uint8)
        self.danger = np.zeros(n, np.uint8)
        self.filtered = np.zeros(n, np.uint8)
        vs = contig.variants_struct()
        P = _ffi.ptr
        oracle_lib().orc_annotate(contig.ref, len(contig.ref), C.byref(vs), int(is_ont), P(self.hom, u8p),
                                  P(self.danger, u8p), P(self.filtered, u8p))


class OraclePhase:
    """Runs the oracle stage by stage on a synth.Contig; results are numpy copies."""

    def __init__(self, contig, params, apply_filter=True, stages=3):
        lib = oracle_lib()
        P = _ffi.ptr
        self.contig = contig
        self.notes = Notes(contig, params.is_ont)
        vs, bs = contig.variants_struct(), contig.batch_struct()
        oc = OrcCalls()
        rc = lib.orc_call_alleles(C.byref(bs), C.byref(vs), P(self.notes.hom, u8p), P(self.notes.danger, u8p),
                                  P(self.notes.filtered, u8p), int(apply_filter and params.is_ont), C.byref(params),
                                  C.byref(oc))
        self.rc = rc
        nr = oc.n_reads
        self.call_off = g(oc.call_off, nr + 1, np.uint64)
        self.calls = g(oc.calls, oc.n_calls, _ffi.CALL_DTYPE)
        self.read_status = g(oc.read_status, nr, np.uint8)
        self.clip_pos = g(oc.clip_pos, oc.n_clips, np.int32)
        self.clip_front = g(oc.clip_front, oc.n_clips, np.int32)
        self.clip_back = g(oc.clip_back, oc.n_clips, np.int32)
        if stages >= 2 and rc == 0:
            og = OrcGraph()
            lib.orc_build_graph(C.byref(bs), C.byref(vs), P(self.notes.danger, u8p), C.byref(oc), C.byref(params),
                                C.byref(og))
            self.n_aln = og.n_aln
            self.aln_read = g(og.aln_read, og.n_aln, np.int32)
            self.aln_off = g(og.aln_off, og.n_aln + 1, np.uint64)
            self.aln_calls = g(og.aln_calls, int(self.aln_off[-1]) if og.n_aln else 0, _ffi.CALL_DTYPE)
            self.cnv = np.stack([g(og.cnv_start, og.n_cnv, np.int32), g(og.cnv_end, og.n_cnv, np.int32)], 1)
            self.n_nodes = og.n_nodes
            self.node_var = g(og.node_var, og.n_nodes, np.int32)
            self.node_type = g(og.node_type, og.n_nodes, np.uint8)
            self.window = og.window
            self.weights = g(og.weights, og.n_nodes * og.window * 4, np.float32).reshape(og.n_nodes, og.window, 4)
            self.n_contrib, self.n_contrib_far, self.n_far_cells = og.n_contrib, og.n_contrib_far, og.n_far_cells
            if stages >= 3:
                so = OrcSolution()
                lib.orc_solve(C.byref(vs), C.byref(og), C.byref(params), C.byref(so))
                nv = so.n_variants
                self.ps_sweep = g(so.ps_sweep, nv, np.int32)
                self.hap_ref_sweep = g(so.hap_ref_sweep, nv, np.int8)
                self.ps = g(so.ps, nv, np.int32)
                self.hap_ref = g(so.hap_ref, nv, np.int8)
                self.read_hp = g(so.read_hp, so.n_aln, np.int8)
                self.hp_counts = g(so.hp_counts, nv * 4, np.int32).reshape(nv, 4)
                lib.orc_solution_free(C.byref(so))
            lib.orc_graph_free(C.byref(og))
        lib.orc_calls_free(C.byref(oc))


def _tap_calls(tc):
    n = tc.n_aln
    off = g(tc.off, n + 1, np.uint64)
    m = int(off[-1]) if n else 0
    return dict(read_idx=g(tc.read_idx, n, np.int32), off=off, pos=g(tc.pos, m, np.int32),
                allele=g(tc.allele, m, np.int32), quality=g(tc.quality, m, np.int32))


def _tap_nodes(tn):
    n = tn.n
    return dict(pos=g(tn.pos, n, np.int32), type=g(tn.type, n, np.int32), ps=g(tn.ps, n, np.int32),
                hap_ref=g(tn.hap_ref, n, np.int32), hap_alt=g(tn.hap_alt, n, np.int32))


class ReferencePhase:
    """Runs the UNMODIFIED reference (oracle/_ref/libref_tap.so) on a synth.Contig."""

    def __init__(self, contig, params, stop_after_calls=False, chr_name="chrS"):
        lib = tap_lib()
        tin = TapIn(chr=chr_name.encode(), ref=contig.ref, ref_len=len(contig.ref), n_var=contig.n_var,
                    var_pos=_ffi.ptr(contig.var_pos, i32p), var_str_off=_ffi.ptr(contig.var_str_off, _ffi.u32p),
                    var_str=contig.var_str, batch=contig.batch_struct(), names=contig.names,
                    name_stride=contig.NAME_STRIDE, stop_after_calls=int(stop_after_calls), p=params)
        out = TapOut()
        self.rc = lib.ref_tap_phase(C.byref(tin), C.byref(out))
        try:
            self.stage_a = _tap_calls(out.stage_a)
            self.stage_b = _tap_calls(out.stage_b)
            self.clip_pos = g(out.clip_pos, out.n_clips, np.int32)
            self.clip_front = g(out.clip_front, out.n_clips, np.int32)
            self.clip_back = g(out.clip_back, out.n_clips, np.int32)
            self.times = dict(get_snp=out.t_get_snp, filter_snp=out.t_filter_snp, clip=out.t_clip,
                              add_edge=out.t_add_edge, sweep=out.t_sweep, read_correction=out.t_read_correction)
            self.n_empty_after_filter = out.n_empty_after_filter
            self.complete = False
            if self.rc == 0 and not stop_after_calls and out.stage_c.off:
                self.complete = True
                self.stage_c = _tap_calls(out.stage_c)
                self.cnv = np.stack([g(out.cnv_start, out.n_cnv, np.int32), g(out.cnv_end, out.n_cnv, np.int32)], 1)
                self.n_edge_nodes, self.n_cells, self.n_contrib = out.n_edge_nodes, out.n_cells, out.n_contrib
                self.cell_a = g(out.cell_a, out.n_cells, np.int32)
                self.cell_b = g(out.cell_b, out.n_cells, np.int32)
                self.cell_which = g(out.cell_which, out.n_cells, np.uint8)
                self.cell_val = g(out.cell_val, out.n_cells, np.float32)
                self.nodes_sweep = _tap_nodes(out.nodes_sweep)
                self.nodes_final = _tap_nodes(out.nodes_final)
                self.read_hp = g(out.read_hp, out.stage_c.n_aln, np.int32)
                n = out.n_result
                self.res_pos, self.res_block = g(out.res_pos, n, np.int32), g(out.res_block, n, np.int32)
                self.res_hap_ref, self.res_hap_alt = g(out.res_hap_ref, n, np.int32), g(out.res_hap_alt, n, np.int32)
        finally:
            lib.ref_tap_phase_free(C.byref(out))


class ReferencePhaseTimed:
    """The UNMODIFIED reference's phase path on a synth.Contig in the tap's timing mode: nothing is flattened or copied out; only
    the seconds spent inside the reference's own calls (get_snp loop, filterSNP, Clip, addEdge, edgeConnectResult, readCorrection,
    exportResult) and the counts come back.  bench.py's reference arm."""

    def __init__(self, contig, params, chr_name="chrS"):
        lib = tap_lib()
        tin = TapIn(chr=chr_name.encode(), ref=contig.ref, ref_len=len(contig.ref), n_var=contig.n_var,
                    var_pos=_ffi.ptr(contig.var_pos, i32p), var_str_off=_ffi.ptr(contig.var_str_off, _ffi.u32p),
                    var_str=contig.var_str, batch=contig.batch_struct(), names=contig.names,
                    name_stride=contig.NAME_STRIDE, stop_after_calls=2, p=params)
        out = TapOut()
        self.rc = lib.ref_tap_phase(C.byref(tin), C.byref(out))
        self.times = dict(get_snp=out.t_get_snp, filter_snp=out.t_filter_snp, clip=out.t_clip, add_edge=out.t_add_edge,
                          sweep=out.t_sweep, read_correction=out.t_read_correction, export=out.t_export)
        self.seconds = float(sum(self.times.values()))
        self.n_calls, self.n_aln, self.n_result = int(out.n_result_calls), int(out.stage_a.n_aln), int(out.n_result)
        lib.ref_tap_phase_free(C.byref(out))


class OracleTag:
    """Germline haplotag oracle on a phased synth.Contig (contig.phased(...))."""

    def __init__(self, contig, tparams):
        lib = oracle_lib()
        self.notes = Notes(contig, False)
        vs, bs = contig.variants_struct(), contig.batch_struct()
        o = OrcTags()
        self.rc = lib.orc_tag_reads(C.byref(bs), C.byref(vs), _ffi.ptr(self.notes.hom, u8p), C.byref(tparams), C.byref(o))
        n = o.n_reads
        self.category = g(o.category, n, np.uint8)
        self.hp = g(o.hp, n, np.int8)
        self.ps, self.pq = g(o.ps, n, np.int32), g(o.pq, n, np.int32)
        self.h1, self.h2 = g(o.h1, n, np.int32), g(o.h2, n, np.int32)
        self.call_off = g(o.call_off, n + 1, np.uint64)
        self.calls = g(o.calls, o.n_calls, _ffi.CALL_DTYPE)
        lib.orc_tags_free(C.byref(o))


class ReferenceTag:
    """The UNMODIFIED reference's germline haplotag objects on a phased synth.Contig (oracle/ref_tap_tag.cpp)."""

    def __init__(self, contig, tparams, chr_name="chrS"):
        lib = tap_lib()
        tin = TapTagIn(chr=chr_name.encode(), ref=contig.ref, ref_len=len(contig.ref), n_var=contig.n_var,
                       var_pos=_ffi.ptr(contig.var_pos, i32p), var_str_off=_ffi.ptr(contig.var_str_off, _ffi.u32p),
                       var_str=contig.var_str, var_hp1_is_alt=_ffi.ptr(contig.var_hp1_is_alt, u8p),
                       var_ps=_ffi.ptr(contig.var_ps, i32p), batch=contig.batch_struct(), names=contig.names,
                       name_stride=contig.NAME_STRIDE, p=tparams)
        out = TapTagOut()
        self.rc = lib.ref_tap_tag(C.byref(tin), C.byref(out))
        n = out.n_reads
        self.category = g(out.category, n, np.uint8)
        for k in ("hp", "ps", "pq", "h1", "h2", "n_ps"):
            setattr(self, k, g(getattr(out, k), n, np.int32))
        self.var_off = g(out.var_off, n + 1, np.uint64)
        m = int(self.var_off[-1]) if n else 0
        self.var_pos, self.var_hp = g(out.var_pos, m, np.int32), g(out.var_hp, m, np.int32)
        self.ps_off = g(out.ps_off, n + 1, np.uint64)
        m = int(self.ps_off[-1]) if n else 0
        self.ps_id, self.ps_count = g(out.ps_id, m, np.int32), g(out.ps_count, m, np.int32)
        self.stats = dict(zip(_ffi.TAG_COUNTERS, list(out.stats)))
        self.t_total = out.t_total
        lib.ref_tap_tag_free(C.byref(out))


def _somatic_fields(self, o):
    n, nt = o.n_reads, o.n_tum
    self.n_reads, self.n_tum = n, nt
    self.tum_var = g(o.tum_var, nt, np.int32)
    self.category = g(o.category, n, np.uint8)
    self.read_hp, self.hp_before = g(o.read_hp, n, np.int8), g(o.hp_before, n, np.int8)
    for k in ("ps", "pq", "h1", "h2", "h3", "end_pos", "read_len"):
        setattr(self, k, g(getattr(o, k), n, np.int32))
    self.n_ps = g(o.n_ps, n, np.uint8)
    self.derive_similarity = g(o.derive_similarity, n, np.float32)
    self.pos_base = g(o.pos_base, nt * 15, np.int32).reshape(nt, 15)
    for k in ("read_hp_count", "somatic_read_hp_count", "hp_before_count", "hp_after_count", "h3_before_count", "h3_after_count"):
        setattr(self, k, g(getattr(o, k), nt * 9, np.int32).reshape(nt, 9))
    self.case_count = g(o.case_count, nt * 6, np.int32).reshape(nt, 6)
    self.allele_count = g(o.allele_count, nt * 2, np.int32).reshape(nt, 2)
    self.window_hist = g(o.window_hist, nt * 2 * 201, np.int32).reshape(nt, 2, 201)
    self.cover_start, self.cover_end = g(o.cover_start, nt, np.int32), g(o.cover_end, nt, np.int32)
    self.ratios_f = g(o.ratios_f, nt * 9, np.float32).reshape(nt, 9)
    self.ratios_d = g(o.ratios_d, nt * 4, np.float64).reshape(nt, 4)
    self.case_read_count = g(o.case_read_count, nt, np.int32)
    self.n_window_items = int(o.n_window_items)
    self.call_off = g(o.call_off, n + 1, np.uint64)
    self.calls = g(o.calls, o.n_calls, _ffi.CALL_DTYPE)


class OracleSomatic:
    """oracle_somatic.c on a union contig (synth.Contig.somatic_union): mode in SOM_MODES."""

    def __init__(self, contig, tparams, mode):
        lib = oracle_lib()
        self.notes = Notes(contig, False)
        vs, bs, ts = contig.variants_struct(), contig.batch_struct(), contig.tumor_struct()
        o = OrcSomaticOut()
        self.rc = lib.orc_somatic(SOM_MODES[mode], C.byref(bs), C.byref(vs), C.byref(ts), _ffi.ptr(self.notes.hom, u8p), contig.ref,
                                  len(contig.ref), C.byref(tparams), C.byref(o))
        _somatic_fields(self, o)
        lib.orc_somatic_free(C.byref(o))


class ReferenceSomatic:
    """The UNMODIFIED reference's extract / somatic-tagging objects on a union contig (oracle/ref_tap_somatic.cpp).
    Per-read values the reference keeps in locals come back as -1 (h1/h2/h3/end_pos/read_len), 255 (n_ps)."""

    @staticmethod
    def tap_input(contig, tparams, chr_name, stats, keep):
        P = _ffi.ptr
        return TapSomIn(chr=chr_name.encode(), ref=contig.ref, ref_len=len(contig.ref), n_var=contig.n_var, var_pos=P(contig.var_pos, i32p),
                        var_str_off=P(contig.var_str_off, _ffi.u32p), var_str=contig.var_str, var_hp1_is_alt=P(contig.var_hp1_is_alt, u8p),
                        var_ps=P(contig.var_ps, i32p), nor_gt=P(contig.var_gt_kind, u8p), nor_present=P(contig.nor_present, u8p),
                        tum_present=P(contig.tum_present, u8p), tum_str_off=P(contig.tum_str_off, _ffi.u32p), tum_str=contig.tum_str,
                        tum_gt=P(contig.tum_gt, u8p), tum_hp1_is_alt=P(contig.tum_hp1_is_alt, u8p), tum_ps=P(contig.tum_ps, i32p),
                        is_somatic=P(contig.is_somatic, u8p), derive_hp=P(contig.derive_hp, i8p), batch=contig.batch_struct(),
                        names=contig.names, name_stride=contig.NAME_STRIDE, p=tparams,
                        stats_out=C.cast(stats, C.POINTER(C.c_int64)), keep=keep)

    def __init__(self, contig, tparams, mode, chr_name="chrS", keep=None):
        lib = tap_lib()
        P = _ffi.ptr
        stats = (C.c_int64 * 22)()
        tin = TapSomIn(chr=chr_name.encode(), ref=contig.ref, ref_len=len(contig.ref), n_var=contig.n_var, var_pos=P(contig.var_pos, i32p),
                       var_str_off=P(contig.var_str_off, _ffi.u32p), var_str=contig.var_str, var_hp1_is_alt=P(contig.var_hp1_is_alt, u8p),
                       var_ps=P(contig.var_ps, i32p), nor_gt=P(contig.var_gt_kind, u8p), nor_present=P(contig.nor_present, u8p),
                       tum_present=P(contig.tum_present, u8p), tum_str_off=P(contig.tum_str_off, _ffi.u32p), tum_str=contig.tum_str,
                       tum_gt=P(contig.tum_gt, u8p), tum_hp1_is_alt=P(contig.tum_hp1_is_alt, u8p), tum_ps=P(contig.tum_ps, i32p),
                       is_somatic=P(contig.is_somatic, u8p), derive_hp=P(contig.derive_hp, i8p), batch=contig.batch_struct(),
                       names=contig.names, name_stride=contig.NAME_STRIDE, p=tparams,
                       stats_out=C.cast(stats, C.POINTER(C.c_int64)), keep=keep)
        o = OrcSomaticOut()
        self.rc = lib.ref_tap_somatic(SOM_MODES[mode], C.byref(tin), C.byref(o))
        _somatic_fields(self, o)
        self.stats = dict(zip(_ffi.SOMATIC_COUNTERS + [f"hp{k}" for k in range(9)], list(stats)))
        lib.ref_tap_somatic_free(C.byref(o))


class ReferencePurity:
    """The UNMODIFIED reference's extract passes followed by its TumorPurityEstimator (oracle/ref_tap_somatic.cpp)."""

    def __init__(self, normal_contig, tumor_contig, tparams, chr_name="chrS"):
        lib = tap_lib()
        st = lib.ref_tap_state_new()
        try:
            self.normal = ReferenceSomatic(normal_contig, tparams, "extract_normal", chr_name, keep=st)
            self.tumor = ReferenceSomatic(tumor_contig, tparams, "extract_tumor", chr_name, keep=st)
            o = TapPurityOut()
            lib.ref_tap_purity(st, chr_name.encode(), C.byref(o))
            self.result = {f: getattr(o, f) for f, _ in o._fields_}
            self.purity = o.purity
        finally:
            lib.ref_tap_state_free(st)


class ReferenceSomaticCall:
    """The UNMODIFIED reference's two extract passes followed by the private stages of SomaticVarCaller::variantCalling and
    getSomaticFlag on its own maps (oracle/ref_tap_somatic.cpp: ref_tap_somatic_call)."""

    def __init__(self, normal_contig, tumor_contig, tparams, purity, enable_filter=True, chr_name="chrS"):
        lib = tap_lib()
        st = lib.ref_tap_state_new()
        try:
            self.normal = ReferenceSomatic(normal_contig, tparams, "extract_normal", chr_name, keep=st)
            self.tumor = ReferenceSomatic(tumor_contig, tparams, "extract_tumor", chr_name, keep=st)
            stats = (C.c_int64 * 22)()
            tin = ReferenceSomatic.tap_input(tumor_contig, tparams, chr_name, stats, None)
            o = TapCallOut()
            self.rc = lib.ref_tap_somatic_call(st, C.byref(tin), float(purity), int(enable_filter), C.byref(o))
            nt, n = o.n_tum, o.n_reads
            self.touched, self.in_dense = g(o.touched, nt, np.uint8), g(o.in_dense, nt, np.uint8)
            self.mean_alt, self.z_score = g(o.mean_alt, nt, np.float32), g(o.z_score, nt, np.float32)
            self.interval_snp_count, self.min_distance = g(o.interval_snp_count, nt, np.int32), g(o.min_distance, nt, np.int32)
            self.dense_alt_same = g(o.dense_alt_same, nt, np.int32)
            self.filtered_by = g(o.filtered_by, nt * 6, np.uint8).reshape(nt, 6)
            self.is_filter_out, self.high_con = g(o.is_filter_out, nt, np.uint8), g(o.high_con, nt, np.uint8)
            self.derive_hp, self.flag_derive_hp = g(o.derive_hp, nt, np.int32), g(o.flag_derive_hp, nt, np.int32)
            self.is_somatic = g(o.is_somatic, nt, np.uint8)
            self.read_hp, self.read_h3 = g(o.read_hp, n, np.int8), g(o.read_h3, n, np.int32)
            self.tier = o.tier
            lib.ref_tap_somatic_call_free(C.byref(o))
        finally:
            lib.ref_tap_state_free(st)
