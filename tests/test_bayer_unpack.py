"""Lytro raw bit-unpacking ("next" row of SURVEY §8f): oracle vs golden vectors of the unmodified reference (CPU),
CUDA unpack vs the same vectors and at full Illum size vs the oracle (GPU)."""
import numpy as np
import pytest

from conftest import golden
from oracle import oracle_np as onp


@pytest.mark.parametrize("bits", [10, 12])
def test_oracle_comp_bayer_equals_reference_golden(bits):
    g = golden("comp_bayer")
    got = onp.comp_bayer(g["buf%d" % bits].tobytes(), tuple(g["shape%d" % bits]), bits)
    assert got.dtype == np.float64 and np.array_equal(got, g["img%d" % bits])


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [10, 12])
def test_comp_bayer_vs_reference_golden(cuda, bits):
    from plenopticam_b200.lfp_reader import comp_bayer
    g = golden("comp_bayer")
    got = comp_bayer(bytearray(g["buf%d" % bits].tobytes()), list(g["shape%d" % bits]), bits)
    assert got.dtype == np.float64 and np.array_equal(got, g["img%d" % bits])
    with pytest.raises(ValueError):
        comp_bayer(g["buf%d" % bits].tobytes()[:-15], list(g["shape%d" % bits]), bits)
    with pytest.raises(ValueError):
        comp_bayer(g["buf%d" % bits].tobytes(), list(g["shape%d" % bits]), 8)


@pytest.mark.gpu
def test_comp_bayer_illum_size_equals_oracle(cuda):
    from plenopticam_b200.lfp_reader import comp_bayer_device
    from plenopticam_b200 import ops
    rng = np.random.default_rng(52)
    shape = (7728, 5368)
    buf = rng.integers(0, 256, size=shape[0] * shape[1] * 10 // 8, dtype=np.uint8)
    got = ops.to_host(comp_bayer_device(buf, shape, 10))
    assert got.shape == (5368, 7728) and got.dtype == np.uint16
    assert np.array_equal(got, onp.comp_bayer(buf.tobytes(), shape, 10).astype(np.uint16))
