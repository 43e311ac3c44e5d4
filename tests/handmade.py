"""Hand-written contigs for edge cases: explicit reference, variants and reads (CIGAR string, sequence, qualities)."""
import importlib
import re

import numpy as np

synth = importlib.import_module("longphase_s_b200.synth")

OPS = "MIDNSHP=XB"
NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


class ManualContig(synth.Contig):
    def __init__(self, ref, variants, reads):
        """variants: [(pos0, REF, ALT, hp1_is_alt)], reads: [dict(name, pos, cigar, seq, qual, flag=0, mapq=60)] sorted by pos."""
        self.ref = ref.encode()
        variants = sorted(variants)
        self.n_var = len(variants)
        self.var_pos = np.array([v[0] for v in variants], np.int32)
        self.var_ref0 = np.array([ord(v[1][0]) for v in variants], np.uint8)
        self.var_alt0 = np.array([ord(v[2][0]) for v in variants], np.uint8)
        self.var_ref_len = np.array([len(v[1]) for v in variants], np.uint16)
        self.var_alt_len = np.array([len(v[2]) for v in variants], np.uint16)
        self.var_hp1_is_alt = np.array([v[3] if len(v) > 3 else 0 for v in variants], np.uint8)
        blob, off = b"", [0]
        for v in variants:
            blob += v[1].encode() + b"\0" + v[2].encode() + b"\0"
            off.append(len(blob))
        self.var_str, self.var_str_off = blob, np.array(off, np.uint32)
        n = len(reads)
        self.n_reads = n
        self.ref_start = np.array([r["pos"] for r in reads], np.int32)
        self.flag = np.array([r.get("flag", 0) for r in reads], np.uint16)
        self.mapq = np.array([r.get("mapq", 60) for r in reads], np.uint8)
        cig, seq4, qual = [], [], []
        self.cigar_off, self.seq_off, self.qual_off = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        self.n_cigar, self.l_qseq = np.zeros(n, np.uint32), np.zeros(n, np.int32)
        co = so = qo = 0
        for i, r in enumerate(reads):
            ops = r["cigar"] if isinstance(r["cigar"], list) else [(int(a), OPS.index(b)) for a, b in re.findall(r"(\d+)([MIDNSHP=XB])", r["cigar"])]
            self.cigar_off[i], self.seq_off[i], self.qual_off[i] = co, so, qo
            self.n_cigar[i] = len(ops)
            cig += [(ln << 4) | op for ln, op in ops]
            s = r.get("seq", "")
            self.l_qseq[i] = len(s)
            codes = [NT16[c] for c in s]
            if len(codes) & 1:
                codes.append(0)
            seq4 += [(codes[k] << 4) | codes[k + 1] for k in range(0, len(codes), 2)]
            q = r.get("qual", [30] * len(s))
            qual += list(q)
            co += len(ops); so += (len(s) + 1) // 2; qo += len(s)
        self.cigar = np.array(cig, np.uint32)
        self.seq4 = np.array(seq4, np.uint8)
        self.qual = np.array(qual, np.uint8)
        names = [r["name"] for r in reads]
        order = {nm: k for k, nm in enumerate(sorted(set(names)))}
        self.name_rank = np.array([order[nm] for nm in names], np.int32)
        self.hap = np.zeros(n, np.uint8)
        self.names = b"".join(nm.encode().ljust(self.NAME_STRIDE, b"\0") for nm in names)


def read_from_ref(ref, pos, cigar, name, edits=None, qual=30, **kw):
    """Builds the read sequence implied by `cigar` over `ref` starting at `pos`; `edits` {query_index: base}."""
    ops = [(int(a), b) for a, b in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
    seq, rp = [], pos
    for ln, op in ops:
        if op in "M=X":
            seq += list(ref[rp:rp + ln]); rp += ln
        elif op in "IS":
            seq += ["A"] * ln
        elif op in "DN":
            rp += ln
    for k, b in (edits or {}).items():
        seq[k] = b
    d = dict(name=name, pos=pos, cigar=cigar, seq="".join(seq), qual=[qual] * len(seq))
    d.update(kw)
    return d


class ManualUnion(ManualContig):
    """A hand-written union variant map (std::map<int, MultiGenomeVar>): entries are dicts
    pos, nor=(REF, ALT, hp1_is_alt, ps[, gt_kind]) | None, tum=(REF, ALT, gt_kind[, ps]) | None, somatic=0|1, derive=0|1|2."""

    def __init__(self, ref, entries, reads):
        entries = sorted(entries, key=lambda e: e["pos"])
        base = []
        for e in entries:
            rec = e.get("nor") or e["tum"]
            base.append((e["pos"], rec[0], rec[1], (e["nor"][2] if e.get("nor") else 0)))
        super().__init__(ref, base, reads)
        n = len(entries)
        self.var_ps = np.array([(e["nor"][3] if e.get("nor") else 0) for e in entries], np.int32)
        self.var_gt_kind = np.array([((e["nor"][4] if len(e["nor"]) > 4 else 1) if e.get("nor") else 0) for e in entries], np.uint8)
        self.nor_present = np.array([1 if e.get("nor") else 0 for e in entries], np.uint8)
        self.tum_present = np.array([1 if e.get("tum") else 0 for e in entries], np.uint8)
        tum = [e.get("tum") or ("N", "N", 0) for e in entries]
        self.tum_ref0 = np.array([ord(t[0][0]) for t in tum], np.uint8)
        self.tum_alt0 = np.array([ord(t[1][0]) for t in tum], np.uint8)
        self.tum_ref_len = np.array([len(t[0]) for t in tum], np.uint16)
        self.tum_alt_len = np.array([len(t[1]) for t in tum], np.uint16)
        self.tum_gt = np.array([t[2] for t in tum], np.uint8)
        self.tum_hp1_is_alt = np.zeros(n, np.uint8)
        self.tum_ps = np.array([(t[3] if len(t) > 3 else -1) for t in tum], np.int32)
        blob, off = b"", [0]
        for t in tum:
            blob += t[0].encode() + b"\0" + t[1].encode() + b"\0"
            off.append(len(blob))
        self.tum_str, self.tum_str_off = blob, np.array(off, np.uint32)
        self.is_somatic = np.array([e.get("somatic", 0) for e in entries], np.uint8)
        self.derive_hp = np.array([e.get("derive", 0) for e in entries], np.int8)
        self.var_is_somatic = self.is_somatic.copy()
