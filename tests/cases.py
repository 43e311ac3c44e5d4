"""Named synthetic contigs used by the parity tests (small enough for the oracle to finish in seconds)."""
import importlib

import __graft_entry__ as entry

entry.load_package()
synth = importlib.import_module("longphase_s_b200.synth")
ffi = importlib.import_module("longphase_s_b200._ffi")

CASES = {
    # name: (synth kwargs, is_ont)
    "snp_only": (dict(seed=11, contig_len=500_000), True),
    "snp_indel": (dict(seed=12, contig_len=600_000, indel_frac=0.1), True),
    "dense_indel_noseq": (dict(seed=13, contig_len=300_000, indel_frac=0.2, variant_rate=1 / 300.0, noseq_frac=0.3,
                               supp_frac=0.2), True),
    "pacbio_like": (dict(seed=14, contig_len=400_000, indel_frac=0.1, sub_rate=0.003, ins_rate=0.002, del_rate=0.002), False),
    "deep_long_reads": (dict(seed=15, contig_len=400_000, depth=60, mean_len=50_000, variant_rate=1 / 300.0,
                             indel_frac=0.1), True),
    "many_supplementary": (dict(seed=16, contig_len=400_000, supp_frac=0.5, indel_frac=0.05), True),
    "short_reads_sparse": (dict(seed=17, contig_len=800_000, depth=8, mean_len=3_000, variant_rate=1 / 2500.0), True),
}

_cache = {}


def get(name):
    if name not in _cache:
        kw, is_ont = CASES[name]
        _cache[name] = (synth.Contig(**kw), ffi.default_phase_params(is_ont))
    return _cache[name]
