"""The C++ host's output files against GOLDEN digests of the reference binary's files (tests/golden/host_cli.json, written by
tests/golden/make_host_golden.py from oracle/_ref/longphase-s): runs where the reference binary is not available, and pins the
file-level parity of `phase`, `haplotag --log` and `somatic_haplotag` to committed values.  The data sets come from fixed seeds;
the normal VCF of the somatic pair is phased by this host (so the phase golden has to hold first).  CPU: the oracle stands in for the
device, as in tests/test_host_cli.py."""
import json
import os

import pytest

from . import host_cli as hc
from . import test_host_cli as tc
from . import test_host_somatic_cli as ts

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_cli.json")


def germline_files(d):
    a = hc.synth.Contig(seed=611, contig_len=220_000, indel_frac=0.12, depth=16.0, mean_len=8_000.0)
    b = hc.synth.Contig(seed=612, contig_len=150_000, indel_frac=0.06, depth=12.0, mean_len=6_000.0, supp_frac=0.25)
    e = hc.synth.Contig(seed=613, contig_len=20_000, depth=4.0, mean_len=3_000.0)
    os.makedirs(os.path.join(d, "germline"), exist_ok=True)
    return hc.write_dataset(os.path.join(d, "germline"), [("chrA", a, True), ("chrEmpty", e, False), ("chrB", b, True)], fast_bam=True)


def somatic_files(d):
    pairs = []
    for name, kw in (("chrA", dict(seed=631, contig_len=300_000, indel_frac=0.15, somatic_rate=1 / 3000.0)),):
        cn = hc.synth.Contig(**kw, depth=25, purity=0.0, read_seed=1000 + kw["seed"])
        ct = hc.synth.Contig(**kw, depth=50, purity=0.6, read_seed=2000 + kw["seed"])
        pairs.append((name, cn, ct))
    os.makedirs(os.path.join(d, "somatic"), exist_ok=True)
    files = hc.write_somatic_dataset(os.path.join(d, "somatic"), pairs)
    # the phased NORMAL VCF: this host's `phase` on the normal BAM (identical to the reference's by the phase golden)
    g = dict(vcf=files["germline_vcf"], bam=files["normal_bam"], fasta=files["fasta"])
    tc.oracle_phase_through_host(g, ["--ont", "--indels"], os.path.join(d, "somatic", "phase"))
    files["normal_vcf"] = os.path.join(d, "somatic", "phase", "out.vcf")
    return files


def PHASE_ARGS(files):
    return tc.phase_args(files, ["--ont", "--indels"])


def TAG_ARGS(files, vcf):
    return tc.tag_args(files, vcf, ["--log", "--tagSupplementary"])


def SOM_ARGS(files):
    return ts.som_args(files, ["--output-somatic-vcf"])


def log_without_paths(text):
    return "\n".join(ln for ln in text.split("\n") if not ln.startswith("##"))


@tc.needs_host
@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden digests missing")
def test_host_files_match_golden_digests(tmp_path):
    gold = json.load(open(GOLDEN))
    d = str(tmp_path)
    files = germline_files(d)
    tc.oracle_phase_through_host(files, ["--ont", "--indels"], os.path.join(d, "own"))
    vcf = os.path.join(d, "own", "out.vcf")
    assert hc.text_digest(hc.strip_commandline(open(vcf).read())) == gold["phase_vcf"], "phased VCF differs from the reference's"
    tc.oracle_tag_pipelined(files, vcf, ["--log", "--tagSupplementary"], os.path.join(d, "own"), 300)
    assert hc.bam_digest(os.path.join(d, "own", "tagged.bam")) == gold["haplotag_bam"], "tagged BAM differs from the reference's"
    assert hc.text_digest(log_without_paths(open(os.path.join(d, "own", "tagged.out")).read())) == gold["haplotag_log"]
    sfiles = somatic_files(d)
    ts.oracle_somatic_through_host(sfiles, ["--output-somatic-vcf"], os.path.join(d, "sown"), chunk=400, pipelined=True)
    assert hc.bam_digest(os.path.join(d, "sown", "som.bam")) == gold["somatic_bam"], "tagged tumor BAM differs from the reference's"
    assert hc.text_digest(open(os.path.join(d, "sown", "som_purity.out")).read()) == gold["somatic_purity_out"]
    assert hc.text_digest(hc.strip_commandline(open(os.path.join(d, "sown", "som_sc.vcf")).read())) == gold["somatic_sc_vcf"]
