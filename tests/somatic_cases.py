"""Tumor / normal pairs for the somatic family: one synthetic contig (same reference and variants), a NORMAL read set
without tumor cells and a TUMOR read set at a given purity, and the union variant map a caller would hold
(synth.Contig.somatic_union)."""
import importlib

from . import cases  # noqa: F401  (loads the package)

synth = importlib.import_module("longphase_s_b200.synth")
ffi = importlib.import_module("longphase_s_b200._ffi")

SOMATIC_CASES = {
    # name: (synth kwargs shared by both read sets, normal depth, tumor depth, purity, union seed)
    "snv_indel": (dict(seed=21, contig_len=400_000, indel_frac=0.15, somatic_rate=1 / 4000.0), 25, 50, 0.6, 5),
    "dense_somatic": (dict(seed=22, contig_len=250_000, indel_frac=0.3, variant_rate=1 / 400.0, somatic_rate=1 / 800.0, supp_frac=0.2,
                           lowq_mapq_frac=0.25), 20, 40, 0.8, 6),
    "low_purity_long": (dict(seed=23, contig_len=500_000, indel_frac=0.1, somatic_rate=1 / 6000.0, mean_len=40_000), 15, 30, 0.3, 7),
}
_cache = {}


def param_sets():
    T = ffi.LpsTagParams
    return {"purity_q20": T(mapping_quality=20, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6),
            "tag_q1": T(mapping_quality=1, mapq_filter=1, tag_supplementary=0, have_reference=1, percentage_threshold=0.6),
            "p90_q30": T(mapping_quality=30, mapq_filter=1, tag_supplementary=1, have_reference=1, percentage_threshold=0.9)}


def get(name):
    """-> (normal_case, tumor_case): same union table, the NORMAL and the TUMOR read batch."""
    if name not in _cache:
        kw, dn, dt, purity, useed = SOMATIC_CASES[name]
        cn = synth.Contig(**kw, depth=dn, purity=0.0, read_seed=1000 + kw["seed"])
        ct = synth.Contig(**kw, depth=dt, purity=purity, read_seed=2000 + kw["seed"])
        un = cn.somatic_union(seed=useed)
        _cache[name] = (un, un.with_reads_of(ct))
    return _cache[name]
