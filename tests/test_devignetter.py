"""De-vignetting (first "next" row of SURVEY §8f): oracle vs golden vectors of the unmodified reference (CPU),
CUDA path vs the same vectors (GPU, through the C-ABI)."""
import numpy as np
import pytest

from conftest import golden, make_cfg
from oracle import oracle_np as onp

NAMES = ["devignetter_rgb_u16", "devignetter_mono_f64"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_devignette_equals_reference_golden(name):
    g = golden(name)
    out, wn, noise = onp.devignette(g["lfp"], g["wht"], float(g["pitch_mean"]))
    assert np.array_equal(out, g["out"]) and out.dtype == np.float64
    assert np.array_equal(wn, g["wht_out"])
    assert abs(noise - float(g["noise_lev"])) < 1e-15
    assert (g["wht"].reshape(-1) == 0).any()                      # the where=wht != 0 branch is exercised


def test_gauss_kernel_is_the_separable_factor_of_the_reference_kernel():
    from plenopticam_b200.lfp_aligner.lfp_devignetter import gauss_kernel_1d
    for length in (7, 9, 14, 15):
        n = (length - 1) // 2 + (length + 1) // 2                  # misc/data_proc.py:34
        xs = np.arange(-n // 2 + 1, n // 2 + 1)
        xx, yy = np.meshgrid(xs, xs)
        k2 = np.exp(-(xx ** 2 + yy ** 2) / 2.0)
        k2 /= k2.sum()
        g = gauss_kernel_1d(length)
        assert g.size == n and n % 2 == 1
        assert np.abs(np.outer(g, g) - k2).max() < 1e-16
        assert np.array_equal(g, onp.gauss_kernel_1d(length))


@pytest.mark.gpu
def test_percentile_f64_matches_numpy(cuda):
    from plenopticam_b200 import ops
    import torch
    rng = np.random.default_rng(8)
    x = rng.standard_normal(50021) * 1e3
    x[17] = 0.0
    x[100:120] = x[99]                                             # ties
    d = torch.from_numpy(x).to(cuda)
    for q in ([99.9], [0.0, 50.0, 100.0], [37.123, 0.005]):
        assert np.array_equal(ops.percentile_f64(d, q).cpu().numpy(), np.percentile(x, q)), q


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_devignetter_class_vs_reference_golden(cuda, name):
    from plenopticam_b200.lfp_aligner import LfpDevignetter
    g = golden(name)
    cfg, sta = make_cfg(g["centroids"], 'rec', int(g["M"]))
    lfp_in, wht_in = g["lfp"].copy(), g["wht"].copy()
    obj = LfpDevignetter(lfp_img=lfp_in, wht_img=wht_in, cfg=cfg, sta=sta)
    assert obj.main() is True
    out, wn = obj.lfp_img, obj.wht_img
    assert out.dtype == np.float64 and out.shape == g["out"].shape
    assert np.array_equal(out, g["out"])                           # IEEE float64 divisions: bit-identical
    assert wn.shape == g["wht_out"].shape and np.array_equal(wn, g["wht_out"])
    assert abs(obj.noise_lev - float(g["noise_lev"])) <= 1e-9 * float(g["noise_lev"])
    assert np.array_equal(lfp_in, g["lfp"]) and np.array_equal(wht_in, g["wht"])     # inputs untouched
    dev = obj.lfp_img_device
    assert dev.is_cuda and np.abs(dev.cpu().numpy() - g["out"]).max() <= 1e-6 * np.abs(g["out"]).max()


@pytest.mark.gpu
def test_devignetter_without_white_image_and_interrupt(cuda):
    from plenopticam_b200.lfp_aligner import LfpDevignetter
    g = golden("devignetter_rgb_u16")
    cfg, sta = make_cfg(g["centroids"], 'rec', int(g["M"]))
    obj = LfpDevignetter(lfp_img=g["lfp"], cfg=cfg, sta=sta)
    assert obj.main() is True
    ref, _, _ = onp.devignette(g["lfp"], None, float(g["pitch_mean"]))
    assert np.array_equal(obj.lfp_img, ref)
    sta.interrupt = True
    assert LfpDevignetter(lfp_img=g["lfp"], wht_img=g["wht"], cfg=cfg, sta=sta).main() is False


@pytest.mark.gpu
def test_aligner_devignettes_then_resamples(cuda, tmp_path):
    """LfpAligner with opt_vign and a white image == reference order of operations (devignette -> resample)."""
    from plenopticam_b200.lfp_aligner import LfpAligner
    g = golden("devignetter_rgb_u16")
    M = int(g["M"])
    cfg, sta = make_cfg(g["centroids"], 'rec', M, tmpdir=tmp_path)
    cfg.params[cfg.opt_vign] = True
    cfg.params[cfg.opt_rota] = False
    obj = LfpAligner(g["lfp"], cfg=cfg, sta=sta, wht_img=g["wht"])
    assert obj.main() is True
    got = obj.lfp_img
    ref = onp.uint16_norm(onp.resample(g["out"], g["centroids"], M, 'rec'))
    assert got.dtype == np.uint16 and got.shape == ref.shape
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 1 and (d > 0).mean() < 0.02
