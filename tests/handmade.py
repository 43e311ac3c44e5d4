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
