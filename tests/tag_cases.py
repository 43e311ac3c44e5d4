"""Phased contigs for the haplotag tests: the variants `phase` phased on a synthetic contig (PS / GT from the
oracle's phase result), and a variant with artificial 50-variant blocks so that reads cross phase sets."""
import importlib

import numpy as np

from . import cases

ffi = importlib.import_module("longphase_s_b200._ffi")

TAG_CASES = ["snp_indel", "dense_indel_noseq", "many_supplementary", "deep_long_reads", "short_reads_sparse"]
_cache = {}


def param_sets():
    return {"default": ffi.default_tag_params(),
            "q20_supp": ffi.LpsTagParams(mapping_quality=20, mapq_filter=1, tag_supplementary=1, have_reference=1, percentage_threshold=0.6),
            "nofilter_noref_p75": ffi.LpsTagParams(mapping_quality=1, mapq_filter=0, tag_supplementary=0, have_reference=0, percentage_threshold=0.75)}


def get(name, kind):
    key = (name, kind)
    if key not in _cache:
        from oracle import pyoracle as po
        c, p = cases.get(name)
        if kind == "phase_result":
            orc = po.OraclePhase(c, p)
            _cache[key] = c.phased(orc.ps, orc.hap_ref == 1)
        else:
            blocks = (np.arange(c.n_var) // 50) * 50
            _cache[key] = c.phased(c.var_pos[blocks] + 1)
    return _cache[key]
