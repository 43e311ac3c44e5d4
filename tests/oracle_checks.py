"""Oracle (liboracle.so) against a reference observation — either live from the tap or from a golden .npz."""
import numpy as np

from . import compare as cmp


class RefView:
    """Uniform view over a live po.ReferencePhase or a golden fixture."""

    def __init__(self, src):
        if isinstance(src, dict) or hasattr(src, "files"):
            g = src
            st = lambda p: dict(read_idx=g[f"{p}_read_idx"], off=g[f"{p}_off"], pos=g[f"{p}_pos"], allele=g[f"{p}_allele"],  # noqa: E731
                                quality=g[f"{p}_quality"])
            nd = lambda p: dict(pos=g[f"{p}_pos"], type=g[f"{p}_type"], ps=g[f"{p}_ps"], hap_ref=g[f"{p}_hap_ref"],  # noqa: E731
                                hap_alt=g[f"{p}_hap_alt"])
            self.stage_a, self.stage_b, self.stage_c = st("a"), st("b"), st("c")
            self.nodes_sweep, self.nodes_final = nd("sweep"), nd("final")
            for k in ("clip_pos", "clip_front", "clip_back", "cnv", "cell_a", "cell_b", "cell_which", "cell_val", "read_hp",
                      "res_pos", "res_block", "res_hap_ref", "res_hap_alt"):
                setattr(self, k, g[k])
            self.n_contrib = int(g["n_contrib"][0])
        else:
            self.__dict__.update(src.__dict__)


def check_oracle_against(ref, contig, params, po):
    ref = RefView(ref)
    orc_a = po.OraclePhase(contig, params, apply_filter=False, stages=1)
    orc = po.OraclePhase(contig, params)
    vp = contig.var_pos
    cmp.assert_same_calls(cmp.calls_by_read(orc_a.call_off, orc_a.calls, vp), cmp.tap_stage_by_read(ref.stage_a), "stage A (get_snp)")
    for k in ("clip_pos", "clip_front", "clip_back"):
        assert np.array_equal(getattr(orc, k), getattr(ref, k)), k
    cmp.assert_same_calls(cmp.calls_by_read(orc.call_off, orc.calls, vp), cmp.tap_stage_by_read(ref.stage_b), "stage B (filterSNP)",
                          allow_empty_in_b=True)
    oc = {}
    for k, r in enumerate(orc.aln_read):
        c = orc.aln_calls[int(orc.aln_off[k]):int(orc.aln_off[k + 1])]
        oc[int(r)] = (vp[c["var"]].astype(np.int64), c["allele"].astype(np.int64), c["quality"].astype(np.int64))
    cmp.assert_same_calls(oc, cmp.tap_stage_by_read(ref.stage_c), "stage C (addEdge filters)", allow_empty_in_b=True)
    assert np.array_equal(orc.cnv, np.asarray(ref.cnv).reshape(-1, 2)), "CNV intervals"
    node_pos = vp[orc.node_var]
    assert np.array_equal(node_pos, ref.nodes_final["pos"]), "node set"
    assert np.array_equal(orc.node_type.astype(np.int32), ref.nodes_final["type"]), "node types"
    tab, far = cmp.dense_from_cells(ref, node_pos, orc.window)
    assert tab.tobytes() == orc.weights.tobytes(), "edge float bit patterns"
    assert far == orc.n_far_cells and ref.n_contrib == orc.n_contrib + orc.n_contrib_far, "contribution accounting"
    assert np.array_equal(orc.ps_sweep[orc.node_var], ref.nodes_sweep["ps"]), "sweep PS"
    m = ref.nodes_sweep["ps"] != 0
    assert np.array_equal(orc.hap_ref_sweep[orc.node_var][m].astype(np.int32), ref.nodes_sweep["hap_ref"][m]), "sweep haplotypes"
    assert np.array_equal(1 - orc.hap_ref_sweep[orc.node_var][m].astype(np.int32), ref.nodes_sweep["hap_alt"][m]), "sweep alt haplotypes"
    assert np.array_equal(orc.ps[orc.node_var], ref.nodes_final["ps"]), "final PS"
    assert np.array_equal(orc.hap_ref[orc.node_var].astype(np.int32), ref.nodes_final["hap_ref"]), "final haplotypes"
    hp = cmp.name_level_hp(orc.read_hp, orc.aln_read, contig.name_rank)
    # the tap lists stage-C alignments including the ones emptied by filterSNP; align on read index
    ref_hp = {int(r): int(h) for r, h in zip(ref.stage_c["read_idx"], ref.read_hp)}
    for r, h in zip(orc.aln_read, hp):
        assert ref_hp[int(r)] == int(h), f"read haplotype of alignment {r}"
    sel = orc.ps != 0
    assert np.array_equal(vp[sel], ref.res_pos) and np.array_equal(orc.ps[sel], ref.res_block), "exportResult positions / PS"
    assert np.array_equal(orc.hap_ref[sel].astype(np.int32), ref.res_hap_ref), "exportResult GT"
    assert np.array_equal(1 - orc.hap_ref[sel].astype(np.int32), ref.res_hap_alt), "exportResult GT (alt)"
    return orc
