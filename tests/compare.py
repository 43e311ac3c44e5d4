"""Comparison helpers shared by the parity tests: oracle <-> reference tap <-> CUDA path."""
import numpy as np


def calls_by_read(call_off, calls, var_pos):
    """CSR calls -> dict read_idx -> (pos[], allele[], quality[]) for reads with >= 1 call."""
    out = {}
    off = call_off.astype(np.int64)
    nz = np.nonzero(off[1:] > off[:-1])[0]
    for r in nz:
        c = calls[off[r]:off[r + 1]]
        out[int(r)] = (var_pos[c["var"]].astype(np.int64), c["allele"].astype(np.int64), c["quality"].astype(np.int64))
    return out


def tap_stage_by_read(stage):
    out = {}
    off = stage["off"].astype(np.int64)
    for k, r in enumerate(stage["read_idx"]):
        s = slice(off[k], off[k + 1])
        out[int(r)] = (stage["pos"][s].astype(np.int64), stage["allele"][s].astype(np.int64),
                       stage["quality"][s].astype(np.int64))
    return out


def assert_same_calls(a, b, what, allow_empty_in_b=False):
    """a, b: dict read -> (pos, allele, quality).  With allow_empty_in_b, reads present in b with zero calls
    (alignments emptied by filterSNP, which the reference keeps) are ignored."""
    if allow_empty_in_b:
        b = {r: v for r, v in b.items() if len(v[0])}
    assert set(a) == set(b), f"{what}: read sets differ: only-a {sorted(set(a) - set(b))[:5]} only-b {sorted(set(b) - set(a))[:5]}"
    for r in a:
        for x, y, nm in zip(a[r], b[r], ("pos", "allele", "quality")):
            assert np.array_equal(x, y), f"{what}: read {r} {nm} differ\n{x}\n{y}"


def dense_from_cells(ref, node_pos, window):
    """Reference sparse float cells -> dense [n_nodes][window][4] table (+ number of cells outside it)."""
    idx = {int(p): k for k, p in enumerate(node_pos)}
    n = len(node_pos)
    tab = np.zeros((n, window, 4), np.float32)
    far = 0
    for a, b, w, val in zip(ref.cell_a, ref.cell_b, ref.cell_which, ref.cell_val):
        ka, kb = idx[int(a)], idx[int(b)]
        d = kb - ka
        if 1 <= d <= window:
            tab[ka, d - 1, w] = val
        else:
            far += 1
    return tab, far


def name_level_hp(read_hp, aln_read, name_rank):
    """readHpMap is keyed by read NAME: the last alignment of a name wins (PhasingGraph.cpp:945,957)."""
    last = {}
    for hp, r in zip(read_hp, aln_read):
        last[int(name_rank[r])] = int(hp)
    return np.array([last[int(name_rank[r])] for r in aln_read], np.int32)


def well_formed_reads(contig):
    """Reads whose SEQ covers the query span of their CIGAR.  For the others (SEQ '*', l_qseq = 0) the reference's
    tag-family parser reads bases past the end of SEQ without a bounds check (HaplotagParsingBam.cpp:595-596), which is
    undefined; they are excluded from tag-family comparisons with the reference."""
    ok = np.ones(contig.n_reads, bool)
    consumes_q = np.array([1, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0], bool)
    for r in range(contig.n_reads):
        c = contig.cigar[int(contig.cigar_off[r]):int(contig.cigar_off[r]) + int(contig.n_cigar[r])]
        qlen = int((c >> 4)[consumes_q[c & 15]].sum())
        ok[r] = qlen <= int(contig.l_qseq[r])
    return ok


def assert_tag_matches_reference(orc, ref, contig):
    """orc: po.OracleTag (or anything with the same fields), ref: po.ReferenceTag."""
    ok = well_formed_reads(contig)
    assert np.array_equal(orc.category, ref.category), "dispatch categories differ"
    for k in ("hp", "ps", "pq", "h1", "h2"):
        a, b = np.asarray(getattr(orc, k)).astype(np.int64), np.asarray(getattr(ref, k)).astype(np.int64)
        bad = np.nonzero((a != b) & ok)[0]
        assert len(bad) == 0, f"{k} differs at reads {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}"
    off = orc.call_off.astype(np.int64)
    for r in np.nonzero(ok)[0]:
        oc = orc.calls[off[r]:off[r + 1]]
        sel = oc["allele"] >= 0
        s = slice(int(ref.var_off[r]), int(ref.var_off[r + 1]))
        assert np.array_equal(contig.var_pos[oc["var"][sel]], ref.var_pos[s]) and np.array_equal(oc["allele"][sel].astype(np.int32), ref.var_hp[s]), f"variantsHP of read {r}"
        s = slice(int(ref.ps_off[r]), int(ref.ps_off[r + 1]))
        u, cnt = np.unique(contig.var_ps[oc["var"]], return_counts=True)
        assert np.array_equal(u, ref.ps_id[s]) and np.array_equal(cnt, ref.ps_count[s]), f"countPS of read {r}"
    return int(ok.sum())
