"""GPU blocks == host instantiation on the hard BC7 surfaces (two-colour blocks, 0 / 255 extremes, near-flat blocks, one
varying channel, the ends of the value range).  The CPU suite already shows the encoder source byte-identical to the
reference on these (tests/test_bc7.py); this file adds the GPU leg.  It was written after the round's GPU minutes were spent
(the five surfaces of tests/test_bc7.py ARE GPU-verified), so it is named to run last under `pytest -x`."""
import numpy as np
import pytest

from tests.test_bc7 import HARD_KINDS, host_blocks, surface


@pytest.mark.gpu
@pytest.mark.parametrize("kind", HARD_KINDS)
def test_gpu_blocks_equal_host_instantiation_hard_surfaces(cuda, kind):
    for (w, h) in ((256, 192), (70, 50)):
        rgba = surface(6, w, h, kind)
        assert np.array_equal(cuda.bc7_compress(rgba), host_blocks(rgba))
