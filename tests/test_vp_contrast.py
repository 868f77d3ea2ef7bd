"""LfpContrast.main of the extractor (lfp_extractor/lfp_contrast.py:43-67,249-263; options off): histogram alignment
against the central view + sRGB over the whole light field.  Goldens from the unmodified reference
(oracle/gen_golden_contrast.py)."""
import numpy as np
import pytest

from conftest import golden, make_cfg
from oracle import oracle_np as onp

CASES = ("plain", "lo_zero")


def _expected(vp):
    c = vp.shape[0] // 2
    return onp.srgb_conv(onp.auto_hist_align(vp, vp[c, c].astype('float64')))


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_reference_golden(name):
    g = golden("vp_contrast")
    out = _expected(g[name + "_vp"])
    assert out.dtype == np.float32 and np.array_equal(out, g[name + "_out"])


def test_lo_zero_case_really_exercises_the_falsy_bound():
    vp = golden("vp_contrast")["lo_zero_vp"]
    c = vp.shape[0] // 2
    lo, _ = onp.hist_percentiles(vp[c, c].astype('float64'))
    assert lo == 0.0 and vp.min() == 0          # Normalizer then falls back to the data minimum (normalizer.py:40)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("on_device", (False, True))
def test_lfp_contrast_vs_reference_golden(cuda, name, on_device):
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_extractor import LfpContrast
    g = golden("vp_contrast")
    vp, ref = g[name + "_vp"], g[name + "_out"]
    cfg, sta = make_cfg(M=vp.shape[0])
    obj = LfpContrast(vp_img_arr=ops.to_device(vp) if on_device else vp, cfg=cfg, sta=sta)
    assert obj.main() is True
    out = obj.vp_img_arr
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == ref.shape
    assert np.abs(out.astype(np.float64) - ref).max() <= 2e-6          # SFU pow(5/12); bar 1e-4 on [0, 1]


@pytest.mark.gpu
def test_lfp_contrast_options_raise_and_interrupt_returns_false(cuda):
    from plenopticam_b200.lfp_extractor import LfpContrast
    vp = golden("vp_contrast")["plain_vp"]
    cfg, sta = make_cfg(M=vp.shape[0])
    cfg.params[cfg.opt_sat_] = True
    with pytest.raises(NotImplementedError):
        LfpContrast(vp_img_arr=vp, cfg=cfg, sta=sta).main()
    cfg.params[cfg.opt_sat_] = False
    sta.interrupt = True
    assert LfpContrast(vp_img_arr=vp, cfg=cfg, sta=sta).main() is False


@pytest.mark.gpu
def test_extractor_vp_img_arr_is_the_srgb_light_field(cuda, tmp_path):
    """LfpExtractor: vp_img_linear = the gather (what LfpRefocuser receives), vp_img_arr = its colour-managed copy
    (lfp_extractor/top_level.py:87-94); default-on stages outside the path are reported, not silently dropped."""
    from plenopticam_b200.lfp_extractor import LfpExtractor
    g = golden("config1_chain")
    M = int(g["M"])
    cfg, sta = make_cfg(g["centroids"], 'rec', M, tmpdir=tmp_path)
    ex = LfpExtractor(g["aligned"], cfg=cfg, sta=sta)
    assert ex.main() is True
    assert np.array_equal(np.asarray(ex.vp_img_linear), g["vp"])
    exp = _expected(g["vp"])
    out = np.asarray(ex.vp_img_arr)
    assert out.dtype == np.float32 and np.abs(out.astype(np.float64) - exp).max() <= 2e-6
    assert ex.skipped == ['LfpColorEqualizer', 'LfpDepth']            # opt_colo / opt_dpth default to True


@pytest.mark.gpu
def test_contrast_full_light_field_properties(cuda):
    """Illum-sized light field (15 x 15 x 434 x 625 x 3 uint16): output in [0, 1], monotone in the input, and equal to
    the float32 path fed the same values."""
    import torch
    from plenopticam_b200 import ops
    gen = torch.Generator(device='cuda').manual_seed(11)
    vp = torch.randint(0, 65536, (15, 15, 434, 625, 3), generator=gen, device='cuda', dtype=torch.int32).to(torch.uint16)
    out = ops.viewpoints_contrast(vp)
    assert out.dtype == torch.float32 and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    sub = vp[3, 4].to(torch.float32)
    lohi = ops.contrast_bounds(vp[7, 7].to(torch.float32).contiguous())
    assert torch.equal(ops.contrast_srgb(sub.contiguous(), lohi, mm=ops.image_minmax(vp)), out[3, 4])
    order = torch.argsort(sub.flatten()[:100000])
    vals = out[3, 4].flatten()[:100000][order]
    assert bool((vals[1:] >= vals[:-1]).all())
