"""INTEGRATION.md, route A (second snippet): the UNMODIFIED reference `LfpAligner` with its `LfpResampler` / `LfpRotator`
swapped for the B200 classes by attribute assignment, its own config / status objects, its own de-vignetting stage in front.
The reference comes from /root/reference (build container) or the byte-identical snapshot oracle/_ref (GPU box)."""
import tempfile
import warnings

import numpy as np
import pytest

from conftest import golden
from oracle import ref_shim

needs_ref = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference sources not present")


def _patched_reference():
    ref_shim.load_reference()
    import plenopticam.lfp_aligner.top_level as ref_top
    import plenopticam_b200.lfp_aligner.lfp_resampler as r
    import plenopticam_b200.lfp_aligner.lfp_rotator as t
    orig = (ref_top.LfpResampler, ref_top.LfpRotator)
    ref_top.LfpResampler = r.LfpResampler            # lfp_aligner/top_level.py:74
    ref_top.LfpRotator = t.LfpRotator                # lfp_aligner/top_level.py:68
    return ref_top, orig


def _restore(ref_top, orig):
    ref_top.LfpResampler, ref_top.LfpRotator = orig


@needs_ref
def test_reference_aligner_routes_into_the_b200_classes_and_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the parity test below")
    g = golden("config1_chain")
    ref_top, orig = _patched_reference()
    try:
        with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cfg, sta = ref_shim.make_cfg_sta(tmp, g["centroids"], 'rec', int(g["M"]))
            cfg.lfpimg = {}
            cfg.params[cfg.opt_vign] = False
            cfg.params[cfg.opt_rota] = False
            obj = ref_top.LfpAligner(lfp_img=g["lfp"], cfg=cfg, sta=sta)
            with pytest.raises(RuntimeError, match="CUDA"):      # our resampler was reached; it refuses to run on the CPU
                obj.main()
    finally:
        _restore(ref_top, orig)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("rota", [False, True])
def test_unmodified_reference_aligner_with_patched_stages_vs_its_own_golden(cuda, rota):
    """Reference LfpAligner (its float64 copies, its LfpDevignetter) -> our LfpRotator / LfpResampler.  Without rotation the
    result must match the golden the unpatched reference produced (tests/golden/config1_chain.npz)."""
    g = golden("config1_chain")
    M = int(g["M"])
    ref_top, orig = _patched_reference()
    try:
        with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cfg, sta = ref_shim.make_cfg_sta(tmp, g["centroids"], 'rec', M, ran_refo=(-1, 1))
            cfg.lfpimg = {}
            cfg.params[cfg.opt_vign] = True
            cfg.params[cfg.opt_rota] = rota
            obj = ref_top.LfpAligner(lfp_img=g["lfp"], cfg=cfg, sta=sta, wht_img=g["wht"])
            assert obj.main() is True
            aligned = np.asarray(obj.lfp_img)
    finally:
        _restore(ref_top, orig)
    assert aligned.dtype == np.uint16 and aligned.shape == g["aligned"].shape
    if not rota:
        d = np.abs(aligned.astype(np.int64) - g["aligned"].astype(np.int64))
        assert d.max() <= 1 and (d > 0).mean() < 0.02
    else:
        # the centroid table was de-rotated in cfg (lfp_aligner/top_level.py:70) and the image still spans the range
        assert aligned.min() == 0 and aligned.max() == 65535
        assert not np.allclose(np.asarray(cfg.calibs[cfg.mic_list])[:, :2], g["centroids"][:, :2])


@needs_ref
@pytest.mark.parametrize("ndim", [3, 2])
def test_reference_scratch_upsample_alternate_cannot_run(ndim):
    """SURVEY 8 row a11': `LfpShiftAndSum.refo_from_scratch_upsample` (lfp_shiftandsum.py:223-304, `pragma: no cover`, the
    alternate main() takes for an aligned image with opt_refi on) has no output to be equal to: its shape line
    (`m, n, P = shape if len(shape) == 3 else shape[0], shape[1], 1`, :232) binds the whole shape tuple to m for a colour
    image (TypeError at :241) and indexes a third axis of a monochrome one (IndexError at :258).  The drop-in class raises
    NotImplementedError for it (test_cabi_and_host.py::test_unsupported_paths_fail_loudly)."""
    ref_shim.load_reference()
    from plenopticam.lfp_refocuser.lfp_shiftandsum import LfpShiftAndSum
    rng = np.random.default_rng(5)
    img = rng.integers(0, 65536, size=(15, 20, 3) if ndim == 3 else (15, 20), dtype=np.uint16)
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg, sta = ref_shim.make_cfg_sta(tmp, [[0, 0, 0, 0]], 'rec', 5, ran_refo=(-1, 2))
        cfg.params[cfg.opt_refi] = True
        obj = LfpShiftAndSum(lfp_img_align=img, cfg=cfg, sta=sta)
        with pytest.raises(TypeError if ndim == 3 else IndexError):
            obj.main()
