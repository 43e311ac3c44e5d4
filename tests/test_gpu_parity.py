"""-m gpu: the CUDA path, called through the C ABI, must be bit-exact with the oracle."""
import pytest

from . import cases, parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.CASES))
def test_phase_matches_oracle(name):
    contig, params = cases.get(name)
    info = parity.check_phase(contig, params)
    assert info["calls"] > 0
