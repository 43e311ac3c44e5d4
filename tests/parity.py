"""CUDA path (through the C ABI) vs the oracle, stage by stage.  Used by the -m gpu tests and by smoke()."""
import importlib

import numpy as np

import __graft_entry__ as entry

entry.load_package()
host = importlib.import_module("longphase_s_b200.host")
ffi = importlib.import_module("longphase_s_b200._ffi")


def oracle_phase_digest(orc, n_reads):
    """workloads.phase_digest of an oracle.pyoracle.OraclePhase result (what bench.py's gate compares the GPU result with)."""
    workloads = importlib.import_module("longphase_s_b200.workloads")
    hp = np.full(n_reads, -2, np.int8)
    hp[orc.aln_read] = orc.read_hp
    return workloads.phase_digest(orc.ps, orc.hap_ref, hp, orc.hp_counts)


def check_phase(contig, params, ctx=None, verbose=False):
    """Runs the whole phase path on the GPU and asserts bit-exact equality with the oracle at every stage.
    Returns a dict of sizes for reporting."""
    from oracle import pyoracle as po
    own = ctx is None
    if own:
        ctx = host.Context(0)
    try:
        orc = po.OraclePhase(contig, params)
        bp = host.BamParser(ctx, contig, params)
        notes = ctx.notes()
        assert np.array_equal(notes["homopolymer"], orc.notes.hom), "homopolymer notes differ"
        assert np.array_equal(notes["is_danger"], orc.notes.danger), "danger notes differ"
        assert np.array_equal(notes["filtered"], orc.notes.filtered), "filterSNP notes differ"
        calls = bp.direct_detect_alleles(contig)
        assert np.array_equal(calls["read_status"], orc.read_status), "read status differs"
        assert np.array_equal(calls["call_off"], orc.call_off), "call offsets differ"
        assert calls["calls"].tobytes() == orc.calls.tobytes(), "allele calls differ"
        for k in ("clip_pos", "clip_front", "clip_back"):
            assert np.array_equal(calls[k], getattr(orc, k)), f"{k} differs"
        g = host.VairiantGraph(ctx, params)
        edges = g.addEdge()
        assert edges["n_nodes"] == orc.n_nodes, (edges["n_nodes"], orc.n_nodes)
        assert np.array_equal(edges["node_var"], orc.node_var), "node set differs"
        assert np.array_equal(edges["node_type"], orc.node_type), "node types differ"
        assert edges["weights"].tobytes() == orc.weights.tobytes(), "edge weights differ (bit pattern)"
        assert edges["n_contrib"] == orc.n_contrib and edges["n_contrib_far"] == orc.n_contrib_far, "contribution counts differ"
        res = g.phasingProcess()
        assert np.array_equal(res["ps_sweep"], orc.ps_sweep), "sweep phase sets differ"
        assert np.array_equal(res["hap_ref_sweep"], orc.hap_ref_sweep), "sweep haplotypes differ"
        assert np.array_equal(res["ps"], orc.ps), "phase sets differ"
        m = orc.ps != 0
        assert np.array_equal(res["hap_ref"][m], orc.hap_ref[m]), "haplotypes differ"
        assert np.array_equal(res["hap_ref"], orc.hap_ref), "per-variant orientation differs"
        assert np.array_equal(res["hp_counts"], orc.hp_counts), "hp x allele counters differ"
        hp = np.full(contig.n_reads, -2, np.int8)
        hp[orc.aln_read] = orc.read_hp
        assert np.array_equal(res["read_hp"], hp), "read haplotypes differ"
        # the one-call entry point must give the same answer
        ctx.submit(contig.batch_struct())
        res2 = ctx.phase_contig(params)
        for k in ("ps", "hap_ref", "read_hp", "hp_counts"):
            assert np.array_equal(res2[k], res[k]), f"lps_phase_contig {k} differs from the staged calls"
        # ... and so must the compact wire format of the CIGAR stream (lps_read_batch.cigar16)
        ctx.submit(contig.batch_struct16())
        res3 = ctx.phase_contig(params)
        for k in ("ps", "hap_ref", "read_hp", "hp_counts"):
            assert np.array_equal(res3[k], res[k]), f"cigar16 submit: {k} differs"
        # ... and the 8-bit wire format (lps_read_batch.cigar8)
        ctx.submit(contig.batch_struct8())
        res4 = ctx.phase_contig(params)
        for k in ("ps", "hap_ref", "read_hp", "hp_counts"):
            assert np.array_equal(res4[k], res[k]), f"cigar8 submit: {k} differs"
        # ... and SEQ + QUAL as interleaved rows (lps_read_batch.sq): the calls themselves (base quality, allele) and the end result
        ctx.submit(contig.batch_struct_sq())
        calls5 = ctx.call_alleles(params, want_host=True)
        assert np.array_equal(calls5["call_off"], calls["call_off"]) and calls5["calls"].tobytes() == calls["calls"].tobytes(), "sq submit: calls differ"
        assert np.array_equal(calls5["read_status"], calls["read_status"]), "sq submit: read status differs"
        res5 = ctx.phase_contig(params)
        for k in ("ps", "hap_ref", "read_hp", "hp_counts"):
            assert np.array_equal(res5[k], res[k]), f"sq submit: {k} differs"
        info = dict(reads=contig.n_reads, variants=contig.n_var, calls=len(orc.calls), nodes=orc.n_nodes,
                    phased=int(m.sum()), contrib=int(orc.n_contrib), lowq_cells=int((orc.weights != np.round(orc.weights)).sum()),
                    stats=ctx.stats())
        if verbose:
            print(info)
        return info
    finally:
        if own:
            ctx.close()
