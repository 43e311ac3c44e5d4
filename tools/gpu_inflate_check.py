"""LPS_GPU_INFLATE=1 in the real binary on the GPU box: `phase` and `haplotag` reading their BAM regions through lps_bgzf_inflate,
compared with the reference binary's files; wall times with and without it.  Small on purpose (seconds)."""
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
    os.makedirs(cwd, exist_ok=True)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return round(time.perf_counter() - t0, 3), p.returncode, p.stderr


def main():
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    d = tempfile.mkdtemp(prefix="gpu_inflate_")
    contigs = [("chr%d" % (k + 1), hc.synth.Contig(seed=950 + k, contig_len=int(mb * 1e6), indel_frac=0.1, depth=30.0), True) for k in range(2)]
    files = hc.write_dataset(d, contigs, fast_bam=True)
    out = {"reads": sum(c.n_reads for _, c, _ in contigs), "bam_bytes": os.path.getsize(files["bam"])}
    del contigs
    t = str(os.cpu_count() or 8)
    phase = ["phase", "-s", files["vcf"], "-b", files["bam"], "-r", files["fasta"], "-o", "out", "-t", t, "--ont", "--indels"]
    gi = dict(os.environ, LPS_GPU_INFLATE="1")
    out["ref_phase_s"] = run([hc.REF_BIN] + phase, os.path.join(d, "ref"))[0]
    s, rc, err = run([hc.HOST_BIN] + phase, os.path.join(d, "gi"), gi)
    out["own_phase_gpu_inflate"] = {"s": s, "rc": rc, "err": err[-300:] if rc else "", "timing": [x for x in err.split("\n") if x.startswith("[timing]")]}
    out["own_phase_htslib_s"] = run([hc.HOST_BIN] + phase, os.path.join(d, "hts"))[0]
    ref_vcf = os.path.join(d, "ref", "out.vcf")
    if rc == 0:
        out["phase_identical"] = hc.strip_commandline(open(ref_vcf).read()) == hc.strip_commandline(open(os.path.join(d, "gi", "out.vcf")).read())
    tag = ["haplotag", "-s", ref_vcf, "-b", files["bam"], "-r", files["fasta"], "-o", "tagged", "-t", t]
    out["ref_haplotag_s"] = run([hc.REF_BIN] + tag, os.path.join(d, "ref"))[0]
    s, rc, err = run([hc.HOST_BIN] + tag, os.path.join(d, "gi"), gi)
    out["own_haplotag_gpu_inflate"] = {"s": s, "rc": rc, "err": err[-300:] if rc else "", "timing": [x for x in err.split("\n") if x.startswith("[timing]")]}
    if rc == 0:
        out["haplotag_identical"] = hc.bam_payload(os.path.join(d, "gi", "tagged.bam")) == hc.bam_payload(os.path.join(d, "ref", "tagged.bam"))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
