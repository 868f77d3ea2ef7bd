"""BASELINE.json configs[0] in miniature: custom-SPC-like rec MLA, RGB exposure + white image, patch size 9,
LfpAligner (de-vignetting + centroid resampling) -> LfpExtractor -> LfpRefocuser, ran_refo = [-1, 1], against the
outputs of the unmodified reference classes chained the same way (oracle/gen_golden_config1.py)."""
import numpy as np
import pytest

from conftest import golden, make_cfg
from oracle import oracle_np as onp


def test_oracle_chain_equals_reference_golden():
    g = golden("config1_chain")
    M = int(g["M"])
    dv, _, _ = onp.devignette(g["lfp"], g["wht"], float(M))
    al = onp.uint16_norm(onp.resample(dv, g["centroids"], M, 'rec'))
    assert np.array_equal(al, g["aligned"])
    vp = onp.compose_viewpoints(al, M)
    assert np.array_equal(vp, g["vp"])
    st = onp.refocuser_postprocess(onp.refocus_int(vp, tuple(int(v) for v in g["ran_refo"]), M))
    assert st.dtype == np.float32 and np.array_equal(st, g["stack"])


@pytest.mark.gpu
def test_stage_classes_chained_vs_reference_golden(cuda, tmp_path):
    from plenopticam_b200.lfp_aligner import LfpAligner
    from plenopticam_b200.lfp_extractor import LfpExtractor
    from plenopticam_b200.lfp_refocuser import LfpRefocuser
    g = golden("config1_chain")
    M = int(g["M"])
    cfg, sta = make_cfg(g["centroids"], 'rec', M, ran_refo=[int(v) for v in g["ran_refo"]], tmpdir=tmp_path)
    cfg.params[cfg.opt_vign] = True
    cfg.params[cfg.opt_rota] = False
    cfg.params[cfg.opt_view] = True
    cfg.params[cfg.opt_refo] = True
    cfg.params[cfg.opt_pflu] = False
    al = LfpAligner(g["lfp"], cfg=cfg, sta=sta, wht_img=g["wht"])
    assert al.main() is True
    aligned = al.lfp_img
    d = np.abs(aligned.astype(np.int64) - g["aligned"].astype(np.int64))
    assert aligned.dtype == np.uint16 and d.max() <= 1 and (d > 0).mean() < 0.02       # bar: 1e-4 of the range
    ex = LfpExtractor(aligned, cfg=cfg, sta=sta)
    assert ex.main() is True
    vp = np.asarray(ex.vp_img_linear)
    assert np.array_equal(vp, onp.compose_viewpoints(aligned, M))                         # gather: bit-exact
    rf = LfpRefocuser(ex.vp_img_linear, cfg=cfg, sta=sta)
    assert rf.main() is True
    stack = rf.refo_stack
    assert stack.dtype == np.float32 and stack.shape == g["stack"].shape
    assert np.abs(stack.astype(np.float64) - g["stack"]).max() <= 1e-4                    # north-star tolerance
    # and exactly the reference's result when fed the reference's own viewpoints (integer path is exact)
    rf2 = LfpRefocuser(g["vp"], cfg=cfg, sta=sta)
    assert rf2.main() is True
    assert np.abs(rf2.refo_stack.astype(np.float64) - g["stack"]).max() <= 2e-6


@pytest.mark.gpu
def test_pipeline_with_white_image_equals_the_stage_classes(cuda, tmp_path):
    """LightFieldPipeline(wht_img=...) = de-vignetting + resampling + gather + stack + colour management in one executor:
    same aligned image as LfpAligner(opt_vign) and same stack as the chained classes, and the reference within 1e-4."""
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import LfpAligner
    from plenopticam_b200.lfp_extractor import LfpExtractor
    from plenopticam_b200.lfp_refocuser import LfpRefocuser
    from plenopticam_b200.pipeline import LightFieldPipeline
    g = golden("config1_chain")
    M = int(g["M"])
    ran = [int(v) for v in g["ran_refo"]]
    cfg, sta = make_cfg(g["centroids"], 'rec', M, ran_refo=ran, tmpdir=tmp_path)
    for k, v in ((cfg.opt_vign, True), (cfg.opt_rota, False), (cfg.opt_view, True), (cfg.opt_refo, True), (cfg.opt_pflu, False)):
        cfg.params[k] = v
    al = LfpAligner(g["lfp"], cfg=cfg, sta=sta, wht_img=g["wht"])
    assert al.main() is True
    ex = LfpExtractor(al.lfp_img, cfg=cfg, sta=sta)
    assert ex.main() is True
    rf = LfpRefocuser(ex.vp_img_linear, cfg=cfg, sta=sta)
    assert rf.main() is True

    pipe = LightFieldPipeline(g["centroids"], g["lfp"].shape, np.uint16, M, 'rec', ran_refo=ran, wht_img=g["wht"])
    out = pipe.run(g["lfp"]).copy()
    assert np.array_equal(ops.to_host(pipe.aligned), al.lfp_img)
    assert np.array_equal(out, rf.refo_stack)
    assert np.abs(out.astype(np.float64) - g["stack"]).max() <= 1e-4
    assert pipe.noise_lev > 0
