"""Writes tests/golden/host_cli.json: digests of what the UNMODIFIED reference binary (oracle/_ref/longphase-s) writes for the
seeded data sets of tests/test_host_golden.py — phased VCF (without ##commandline), tagged BAM (without the @PG line that carries
the run's paths), --log table, tagged tumor BAM and _purity.out.  Run in the build container:  python tests/golden/make_host_golden.py
The data sets are regenerated from their seeds by the test, so only the digests are stored."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import host_cli as hc  # noqa: E402
from tests import test_host_golden as tg  # noqa: E402
from tests.test_host_cli import run_in  # noqa: E402

if __name__ == "__main__":
    out = {}
    with tempfile.TemporaryDirectory() as d:
        files = tg.germline_files(d)
        run_in(os.path.join(d, "ref"), [hc.REF_BIN] + tg.PHASE_ARGS(files))
        vcf = os.path.join(d, "ref", "out.vcf")
        out["phase_vcf"] = hc.text_digest(hc.strip_commandline(open(vcf).read()))
        run_in(os.path.join(d, "ref"), [hc.REF_BIN] + tg.TAG_ARGS(files, vcf))
        out["haplotag_bam"] = hc.bam_digest(os.path.join(d, "ref", "tagged.bam"))
        out["haplotag_log"] = hc.text_digest(tg.log_without_paths(open(os.path.join(d, "ref", "tagged.out")).read()))
        sfiles = tg.somatic_files(d)
        run_in(os.path.join(d, "sref"), [hc.REF_BIN] + tg.SOM_ARGS(sfiles))
        out["somatic_bam"] = hc.bam_digest(os.path.join(d, "sref", "som.bam"))
        out["somatic_purity_out"] = hc.text_digest(open(os.path.join(d, "sref", "som_purity.out")).read())
        out["somatic_sc_vcf"] = hc.text_digest(hc.strip_commandline(open(os.path.join(d, "sref", "som_sc.vcf")).read()))
    path = os.path.join(ROOT, "tests", "golden", "host_cli.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, out)
