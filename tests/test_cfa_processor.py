"""CfaProcessor.safe_bayer_awb (lfp_aligner/cfa_processor.py:233-255): oracle and CUDA path against goldens produced
by the unmodified reference (oracle/gen_golden_cfa.py)."""
import os

import numpy as np
import pytest

from oracle import oracle_np as onp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "cfa_awb.npz")
CASES = (("grbg_wht_exp", "GRBG"), ("bggr_nowht", "BGGR"), ("grbg_wht0_exp", "GRBG"), ("bggr_wht0", "BGGR"))


def _case(g, name):
    wht = g[name + "_wht"] if name + "_wht" in g.files else None
    exp = float(g[name + "_exp"][0])
    return g[name + "_bay"], wht, [float(v) for v in g[name + "_awb"]], (None if np.isnan(exp) else exp), g[name + "_out"]


@pytest.mark.parametrize("name,pat", CASES)
def test_oracle_equals_reference_golden(name, pat):
    bay, wht, awb, exp, ref = _case(np.load(GOLDEN), name)
    out = onp.cfa_safe_bayer_awb(bay, wht, awb, pat, exp)
    assert out.dtype == ref.dtype == np.float64 and out.shape == ref.shape
    assert np.array_equal(out, ref)


def test_oracle_gain_order_and_unknown_pattern():
    assert onp.cfa_gains([1, 2, 3, 4], "GRBG").tolist() == [3, 2, 1, 4]
    assert onp.cfa_gains([1, 2, 3, 4], "BGGR").tolist() == [1, 3, 4, 2]
    with pytest.raises(ValueError):
        onp.cfa_gains([1, 2, 3, 4], "RGGB")


class _Cfg(object):
    def __init__(self, lfpimg):
        self.lfpimg = lfpimg
        self.params = {'opt_prnt': False}
        self.opt_prnt = 'opt_prnt'


@pytest.mark.gpu
@pytest.mark.parametrize("name,pat", CASES)
def test_cfa_processor_vs_reference_golden(name, pat):
    from plenopticam_b200.lfp_aligner import CfaProcessor
    bay, wht, awb, exp, ref = _case(np.load(GOLDEN), name)
    lfpimg = {'bit': 10, 'bay': pat, 'awb': awb}
    if exp is not None:
        lfpimg['exp'] = exp
    obj = CfaProcessor(bay_img=bay, wht_img=wht, cfg=_Cfg(lfpimg))
    assert obj.safe_bayer_awb() is True
    out = obj.bay_img
    assert out.dtype == np.float64 and out.shape == ref.shape
    if exp is None:
        assert np.array_equal(out, ref)                       # plain IEEE operations only: bit-identical
    else:
        # exp / log of the soft clipping come from CUDA's libm (<= 1 ulp each): 1e-14 relative, far inside 1e-4
        assert np.max(np.abs(out - ref)) <= 1e-14 * max(1.0, float(np.abs(ref).max()))
    with pytest.raises(NotImplementedError):
        obj.main()


@pytest.mark.gpu
def test_cfa_processor_rejects_what_the_reference_cannot_process():
    from plenopticam_b200.lfp_aligner import CfaProcessor
    bay = np.zeros((8, 8), dtype=np.float32)
    with pytest.raises(ValueError):
        CfaProcessor(bay_img=bay, cfg=_Cfg({'bay': 'RGGB', 'awb': [1, 1, 1, 1]})).safe_bayer_awb()
    with pytest.raises(ValueError):
        CfaProcessor(bay_img=np.zeros((8, 8, 3), dtype=np.float32), cfg=_Cfg({'bay': 'GRBG', 'awb': [1, 1, 1, 1]})).safe_bayer_awb()
    with pytest.raises(ValueError):
        CfaProcessor(bay_img=np.zeros((7, 8), dtype=np.float32), cfg=_Cfg({'bay': 'GRBG', 'awb': [1, 1, 1, 1]})).safe_bayer_awb()


@pytest.mark.gpu
def test_cfa_full_sensor_size_properties():
    """Illum sensor size (5368 x 7728 mosaic): unit gains, no white image, no samples above 0.99 -> identity;
    and a uniform gain scales exactly."""
    import torch
    from plenopticam_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(5)
    bay = torch.rand((5368, 7728), generator=g, device='cuda', dtype=torch.float32) * 0.98
    out = ops.cfa_safe_awb(bay, None, [1, 1, 1, 1])
    assert torch.equal(out, bay.double())
    out2 = ops.cfa_safe_awb(bay, None, [2, 2, 2, 2])
    assert torch.equal(out2, bay.double() * 2)
