"""-m gpu: the CUDA path, called through the C ABI, must be bit-exact with the oracle."""
import pytest

from . import cases, parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.CASES))
def test_phase_matches_oracle(name):
    contig, params = cases.get(name)
    info = parity.check_phase(contig, params)
    assert info["calls"] > 0


def test_compact_cigar_with_long_ops():
    """Ops of 4095 bases and more travel in the side table of the compact CIGAR format: same calls, same phasing, same oracle."""
    import copy

    import numpy as np
    contig, params = cases.get("snp_indel")
    c = copy.copy(contig)
    c.cigar = contig.cigar.copy()
    rng = np.random.default_rng(3)
    dels = np.flatnonzero((c.cigar & 15) == 2)
    pick = rng.choice(dels, 60, replace=False)
    c.cigar[pick] = ((4095 + rng.integers(0, 3000, 60)).astype(np.uint32) << 4) | 2
    c.cigar[pick[0]] = (4095 << 4) | 2                      # the smallest length that needs the table
    c.cigar[pick[1]] = (4094 << 4) | 2                      # the largest that does not
    c16, long_len, long_at = c.pack_cigar16()
    assert len(long_len) == 59 and np.array_equal(np.sort(pick[np.arange(60) != 1]), long_at.astype(np.int64))
    info = parity.check_phase(c, params)                    # oracle == uint32 submit == cigar16 submit
    assert info["calls"] > 0
    ctx_a, ctx_b = parity.host.Context(0), parity.host.Context(0)
    try:
        outs = []
        for ctx, batch in ((ctx_a, c.batch_struct()), (ctx_b, c.batch_struct16())):
            ctx.set_reference(c.ref)
            vs = c.variants_struct()
            ctx.set_variants(vs, True)
            ctx.submit(batch)
            outs.append(ctx.call_alleles(params, want_host=True))
        for k in ("call_off", "read_status", "clip_pos", "clip_front", "clip_back"):
            assert np.array_equal(outs[0][k], outs[1][k]), k
        assert outs[0]["calls"].tobytes() == outs[1]["calls"].tobytes()
    finally:
        ctx_a.close(); ctx_b.close()
