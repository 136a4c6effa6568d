"""A short randomised parity sweep in the GPU suite (tools/fuzz_parity.py runs the same loop for minutes): random batch
shapes, channel counts across every engine (tensor cores, zero-padded onto them, warp-per-point, tile), per-axis strides,
voxel sizes, quantised and duplicated points -- count tables bit-exact, sums inside the operator's tolerance."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_cases_match_the_oracle(seed):
    import fuzz_parity
    cases, worst = fuzz_parity.run(6.0, seed)
    assert cases >= 20 and worst < 1e-5
