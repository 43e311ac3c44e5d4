"""File-level end-to-end run: the reference binary (oracle/_ref/longphase-s, CPU, -t all cores) and this repository's binary
(longphase-s_b200/longphase-s-b200, GPU hot path behind the C ABI) on the SAME synthetic BAM / VCF / FASTA, `phase` then `haplotag`.
Checks that the phased VCF (minus ##commandline), the tagged BAM's uncompressed bytes and the --log table are identical, and
prints one JSON object with the wall times.  Used on the GPU box:  python tools/cli_e2e.py --contigs 2 --mb 16 > gpurun_out/cli_e2e.json
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import host_cli as hc  # noqa: E402


def timed(cmd, cwd, env=None):
    os.makedirs(cwd, exist_ok=True)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{cmd} failed: {p.stderr[-1500:]}")
    return dt, p.stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=2)
    ap.add_argument("--mb", type=float, default=16.0)
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--keep", default="")
    ap.add_argument("--variants", action="store_true", help="also time the opt-in paths: LPS_GPU_INFLATE=1, LPS_BAM_LEVEL=1, both")
    ap.add_argument("--deflate", action="store_true", help="also time haplotag with the tagged BAM deflated on the device (LPS_GPU_DEFLATE=1)")
    a = ap.parse_args()
    d = a.keep or tempfile.mkdtemp(prefix="cli_e2e_")
    t0 = time.perf_counter()
    contigs = [("chr%d" % (k + 1), hc.synth.Contig(seed=900 + k, contig_len=int(a.mb * 1e6), indel_frac=0.1, depth=a.depth), True)
               for k in range(a.contigs)]
    files = hc.write_dataset(d, contigs, fast_bam=True)
    n_reads = sum(c.n_reads for _, c, _ in contigs)
    out = {"config": {"contigs": a.contigs, "contig_mb": a.mb, "depth": a.depth, "reads": n_reads, "variants": sum(c.n_var for _, c, _ in contigs),
                      "bam_bytes": os.path.getsize(files["bam"]), "threads": a.threads, "make_dataset_s": round(time.perf_counter() - t0, 2)}}
    del contigs
    t = str(a.threads)
    # the I/O floor: every record of the BAM through htslib's sam_read1 with the same number of BGZF threads, nothing else
    import ctypes
    lib = ctypes.CDLL(hc.HOST_LIB)
    lib.lpsh_decode_only.restype = ctypes.c_int64
    lib.lpsh_decode_only.argtypes = [ctypes.c_char_p, ctypes.c_int]
    t_dec = time.perf_counter()
    n_dec = lib.lpsh_decode_only(files["bam"].encode(), a.threads)
    out["sam_read1_only"] = {"s": round(time.perf_counter() - t_dec, 3), "records": int(n_dec)}
    phase = ["phase", "-s", files["vcf"], "-b", files["bam"], "-r", files["fasta"], "-o", "out", "-t", t, "--ont", "--indels"]
    ref_s, _ = timed([hc.REF_BIN] + phase, os.path.join(d, "ref"))
    own_s, own_err = timed([hc.HOST_BIN] + phase, os.path.join(d, "own"))
    own2_s, _ = timed([hc.HOST_BIN] + phase, os.path.join(d, "own"))       # second run: page cache and CUDA context creation warm
    same = hc.strip_commandline(open(os.path.join(d, "ref", "out.vcf")).read()) == hc.strip_commandline(open(os.path.join(d, "own", "out.vcf")).read())
    out["phase"] = {"reference_s": round(ref_s, 3), "own_s": round(own_s, 3), "own_second_run_s": round(own2_s, 3), "identical_vcf": same,
                    "reads_per_s_reference": n_reads / ref_s, "reads_per_s_own": n_reads / min(own_s, own2_s)}
    vcf = os.path.join(d, "ref", "out.vcf")
    tag = ["haplotag", "-s", vcf, "-b", files["bam"], "-r", files["fasta"], "-o", "tagged", "-t", t, "--log"]
    ref_s, _ = timed([hc.REF_BIN] + tag, os.path.join(d, "ref"))
    own_s, tag_err = timed([hc.HOST_BIN] + tag, os.path.join(d, "own"))
    same_bam = hc.bam_payload(os.path.join(d, "ref", "tagged.bam")) == hc.bam_payload(os.path.join(d, "own", "tagged.bam"))
    same_log = open(os.path.join(d, "ref", "tagged.out")).read() == open(os.path.join(d, "own", "tagged.out")).read()
    out["haplotag"] = {"reference_s": round(ref_s, 3), "own_s": round(own_s, 3), "identical_bam": same_bam, "identical_log": same_log,
                       "reads_per_s_reference": n_reads / ref_s, "reads_per_s_own": n_reads / own_s}
    if a.variants:
        out["variants"] = {}
        for name, extra in (("gpu_inflate", {"LPS_GPU_INFLATE": "1"}), ("bam_level_1", {"LPS_BAM_LEVEL": "1"}),
                            ("gpu_inflate_bam_level_1", {"LPS_GPU_INFLATE": "1", "LPS_BAM_LEVEL": "1"}),
                            ("bam_level_1_readers_4", {"LPS_BAM_LEVEL": "1", "LPS_TAG_READERS": "4"})):
            env = dict(os.environ, **extra)
            v = {}
            if "LPS_GPU_INFLATE" in extra:
                s_phase, err_p = timed([hc.HOST_BIN] + phase, os.path.join(d, name), env)
                v["phase_s"] = round(s_phase, 3)
                v["phase_identical"] = hc.strip_commandline(open(os.path.join(d, name, "out.vcf")).read()) == hc.strip_commandline(open(vcf).read())
            s_tag, err_t = timed([hc.HOST_BIN] + tag, os.path.join(d, name), env)
            v["haplotag_s"] = round(s_tag, 3)
            v["haplotag_identical"] = hc.bam_payload(os.path.join(d, name, "tagged.bam")) == hc.bam_payload(os.path.join(d, "ref", "tagged.bam"))
            v["timing"] = [ln for ln in err_t.split("\n") if ln.startswith("[timing]")]
            out["variants"][name] = v
    if a.deflate:
        out["deflate"] = {"reference_bam_bytes": os.path.getsize(os.path.join(d, "ref", "tagged.bam"))}
        tag_nolog = [w for w in tag if w != "--log"]
        ref2_s, _ = timed([hc.REF_BIN] + tag_nolog, os.path.join(d, "ref_nolog"))
        out["deflate"]["reference_no_log_s"] = round(ref2_s, 3)
        for name, extra in (("htslib_writer", {"LPS_GPU_DEFLATE": "0"}), ("gpu_deflate", {"LPS_GPU_DEFLATE": "1"}), ("gpu_deflate_readers_4", {"LPS_GPU_DEFLATE": "1", "LPS_TAG_READERS": "4"})):
            env = dict(os.environ, **extra)
            s_tag, err_t = timed([hc.HOST_BIN] + tag_nolog, os.path.join(d, name), env)
            v = {"haplotag_s": round(s_tag, 3), "bam_bytes": os.path.getsize(os.path.join(d, name, "tagged.bam")),
                 "haplotag_identical": hc.bam_payload(os.path.join(d, name, "tagged.bam")) == hc.bam_payload(os.path.join(d, "ref_nolog", "tagged.bam")),
                 "timing": [ln for ln in err_t.split("\n") if ln.startswith("[timing]")]}
            if extra.get("LPS_GPU_DEFLATE") == "1":
                v["htslib_reads_it"] = int(lib.lpsh_decode_only(os.path.join(d, name, "tagged.bam").encode(), a.threads)) == int(n_dec)
            out["deflate"][name] = v
    out["own_timing_lines"] = [ln for ln in (own_err + tag_err).split("\n") if ln.startswith("[timing]") or ln.startswith("tag read") or ln.startswith("parsing total")]
    print(json.dumps(out))
    return 0 if (same and same_bam and same_log) else 1


if __name__ == "__main__":
    sys.exit(main())
