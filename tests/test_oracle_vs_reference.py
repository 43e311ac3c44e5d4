"""The oracle against the live reference tap (oracle/_ref/libref_tap.so).  Skipped where the tap was not built."""
import importlib

import numpy as np
import pytest

from . import cases, oracle_checks

po = pytest.importorskip("oracle.pyoracle")
pytestmark = pytest.mark.skipif(not po.tap_available(), reason="oracle/_ref/libref_tap.so not built (needs /root/reference)")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_oracle_matches_live_reference(name):
    contig, params = cases.get(name)
    ref = po.ReferencePhase(contig, params)
    assert ref.rc == 0 and ref.complete
    oracle_checks.check_oracle_against(ref, contig, params, po)


def test_homopolymer_length_matches_reference():
    synth = importlib.import_module("longphase_s_b200.synth")
    ffi = importlib.import_module("longphase_s_b200._ffi")
    contig = synth.Contig(seed=21, contig_len=200_000, indel_frac=0.2, variant_rate=1 / 200.0)
    notes = po.Notes(contig, True)
    lib = po.tap_lib()
    for i in range(0, contig.n_var, 3):
        assert lib.ref_tap_homopolymer(contig.ref, len(contig.ref), int(contig.var_pos[i])) == int(notes.hom[i])
    assert notes.danger.sum() > 0 and notes.filtered.sum() > 0
    assert ffi.default_phase_params(True).connect_adjacent == 35


@pytest.mark.parametrize("q", [0, 1, 20, 61])
def test_mapping_quality_threshold(q):
    contig, params = cases.get("snp_only")
    import copy
    p2 = copy.copy(params)
    p2.mapping_quality = q
    ref = po.ReferencePhase(contig, p2, stop_after_calls=True)
    orc = po.OraclePhase(contig, p2, apply_filter=False, stages=1)
    from . import compare as cmp
    cmp.assert_same_calls(cmp.calls_by_read(orc.call_off, orc.calls, contig.var_pos), cmp.tap_stage_by_read(ref.stage_a), f"q={q}")
    assert np.array_equal(orc.clip_pos, ref.clip_pos)
