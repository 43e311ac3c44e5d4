"""Stage timing of `longphase-s-b200 haplotag` on the GPU box: where the wall time of the tagging pass goes (reader thread, device judge
split into context / tables / submit / kernels, tagging + BAM writing).  Prints the [timing] lines of a few runs as JSON."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import host_cli as hc  # noqa: E402


def run(cmd, cwd, env=None):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return round(time.perf_counter() - t0, 3), p.returncode, [ln for ln in p.stderr.split("\n") if ln.startswith("[timing]")]


def main():
    d = tempfile.mkdtemp(prefix="tag_timing_")
    contigs = [("chr%d" % (k + 1), hc.synth.Contig(seed=900 + k, contig_len=8_000_000, indel_frac=0.1, depth=30.0), True) for k in range(2)]
    files = hc.write_dataset(d, contigs, fast_bam=True)
    n_reads = sum(c.n_reads for _, c, _ in contigs)
    del contigs
    t = str(os.cpu_count() or 8)
    base = ["-b", files["bam"], "-r", files["fasta"], "-t", t]
    out = {"reads": n_reads, "bam_bytes": os.path.getsize(files["bam"]), "threads": int(t)}
    out["reference_phase_s"] = run([hc.REF_BIN, "phase", "-s", files["vcf"], "-o", "out", "--ont", "--indels"] + base, d)[0]
    vcf = os.path.join(d, "out.vcf")
    tag = ["haplotag", "-s", vcf, "-o", "tagged"] + base
    out["reference_haplotag_s"] = run([hc.REF_BIN] + tag, d)[0]
    for name, env in (("default", None), ("chunk_65536", dict(os.environ, LPS_TAG_CHUNK="65536")), ("chunk_2048", dict(os.environ, LPS_TAG_CHUNK="2048"))):
        s, rc, lines = run([hc.HOST_BIN] + tag, d, env)
        out["own_" + name] = {"s": s, "rc": rc, "timing": lines}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
