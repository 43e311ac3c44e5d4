"""__graft_entry__.smoke(): one small contig through every dialect of the CUDA hot path on cuda:0 (phase, germline haplotag,
extract-normal, extract-tumor with the window diff, somatic tagging), each checked bit for bit against the oracle."""
import importlib

import numpy as np

import __graft_entry__ as entry


def run_smoke():
    entry.load_package()
    synth = importlib.import_module("longphase_s_b200.synth")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    host = importlib.import_module("longphase_s_b200.host")
    from oracle import pyoracle as po
    from . import parity
    contig = synth.Contig(seed=5, contig_len=300_000, indel_frac=0.1)
    info = parity.check_phase(contig, ffi.default_phase_params(True))
    print("smoke phase ok:", {k: v for k, v in info.items() if k != "stats"}, "kernel launches", info["stats"]["kernel_launches"])

    ctx = host.Context(0)
    # germline haplotag on the variants `phase` just phased
    orc_p = po.OraclePhase(contig, ffi.default_phase_params(True))
    phased = contig.phased(orc_p.ps, orc_p.hap_ref == 1)
    tp = ffi.default_tag_params()
    res = host.GermlineHaplotagChrProcessor(ctx, phased, tp).processSingleChrom(phased)
    orc = po.OracleTag(phased, tp)
    for k in ("hp", "ps", "pq", "h1", "h2", "category"):
        assert np.array_equal(res[k], getattr(orc, k)), f"haplotag {k} differs from the oracle"
    print("smoke haplotag ok:", int((res["hp"] != 0).sum()), "of", phased.n_reads, "alignments tagged")

    # somatic family on a tumor / normal pair
    kw = dict(seed=6, contig_len=200_000, indel_frac=0.15, somatic_rate=1 / 3000.0)
    cn = synth.Contig(**kw, depth=20, purity=0.0, read_seed=61)
    ct = synth.Contig(**kw, depth=40, purity=0.6, read_seed=62)
    un = cn.somatic_union(seed=3)
    ut = un.with_reads_of(ct)
    sp = ffi.LpsTagParams(mapping_quality=20, mapq_filter=0, tag_supplementary=1, have_reference=1, percentage_threshold=0.6)
    for mode, c, cls, keys in (("extract_normal", un, host.ExtractNorDataChrProcessor, ("pos_base", "read_hp_count", "read_hp")),
                               ("extract_tumor", ut, host.ExtractTumDataChrProcessor,
                                ("pos_base", "read_hp_count", "somatic_read_hp_count", "case_count", "window_hist", "read_hp")),
                               ("somatic_tag", ut, host.SomaticHaplotagChrProcessor,
                                ("hp_before_count", "hp_after_count", "h3_after_count", "cover_start", "cover_end", "read_hp", "ps", "pq"))):
        r = cls(ctx, c, sp).processSingleChrom(c)
        o = po.OracleSomatic(c, sp, mode)
        for k in keys:
            assert np.array_equal(r[k], getattr(o, k)), f"{mode} {k} differs from the oracle"
        print(f"smoke {mode} ok: {c.n_reads} alignments, {r['n_tum']} tumor positions")
    ctx.close()
