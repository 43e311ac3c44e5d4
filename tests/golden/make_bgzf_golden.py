"""Writes tests/golden/htslib_small.bam: a small BAM produced by the reference's OWN htslib (oracle/_ref/mkbam = sam_open "wb" +
sam_write1 of the vendored htslib 1.16, zlib deflate), the golden input of the BGZF inflation tests.  Run in the build container:
    python tests/golden/make_bgzf_golden.py
The expected bytes are not stored: the tests inflate the same file with Python's zlib (gzip members)."""
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
synth = importlib.import_module("longphase_s_b200.synth")

OPS = "MIDNSHP=XB"
NT16 = "=ACMGRSVTWYHKDBN"


def write_sam(contig, path, chrom="chr1"):
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:%s\tLN:%d\n" % (chrom, len(contig.ref)))
        for r in range(contig.n_reads):
            co, nc = int(contig.cigar_off[r]), int(contig.n_cigar[r])
            cigar = "".join("%d%s" % (w >> 4, OPS[w & 15]) for w in contig.cigar[co:co + nc].tolist())
            lq = int(contig.l_qseq[r])
            so, qo = int(contig.seq_off[r]), int(contig.qual_off[r])
            packed = contig.seq4[so:so + (lq + 1) // 2]
            codes = np.stack([packed >> 4, packed & 15], 1).reshape(-1)[:lq]
            seq = "".join(NT16[c] for c in codes.tolist()) or "*"
            qual = "".join(chr(33 + min(int(q), 93)) for q in contig.qual[qo:qo + lq].tolist()) or "*"
            f.write("\t".join([contig.name(r), str(int(contig.flag[r])), chrom, str(int(contig.ref_start[r]) + 1), str(int(contig.mapq[r])),
                               cigar or "*", "*", "0", "0", seq, qual]) + "\n")


if __name__ == "__main__":
    c = synth.Contig(seed=77, contig_len=40_000, depth=3.0, mean_len=5_000.0)
    out = os.path.join(ROOT, "tests", "golden", "htslib_small.bam")
    with tempfile.TemporaryDirectory() as d:
        sam = os.path.join(d, "small.sam")
        write_sam(c, sam)
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "mkbam"), sam, os.path.join(d, "small.bam")])
        os.replace(os.path.join(d, "small.bam"), out)
    print(out, os.path.getsize(out), "bytes,", c.n_reads, "reads")
