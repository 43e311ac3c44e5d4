"""The oracle against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py).
Runs anywhere (no /root/reference, no GPU)."""
import glob
import os

import numpy as np
import pytest

from . import cases, oracle_checks

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[len("phase_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "phase_*.npz")))


def test_fixtures_exist():
    assert len(NAMES) >= 5


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name):
    from oracle import pyoracle as po
    g = np.load(os.path.join(GOLDEN_DIR, f"phase_{name}.npz"))
    contig, params = cases.get(name)
    fp = np.array([contig.n_reads, contig.n_var, int(contig.cigar.sum() % (1 << 31)),
                   int(contig.qual.astype(np.uint64).sum() % (1 << 31)), int(contig.var_pos.astype(np.int64).sum() % (1 << 31))], np.int64)
    assert np.array_equal(fp, g["input_fingerprint"]), "synthetic generator no longer reproduces the fixture's inputs"
    orc = oracle_checks.check_oracle_against(g, contig, params, po)
    assert len(orc.calls) > 0 and (orc.ps != 0).sum() > 0
