"""BASELINE.json configs[0] in miniature, through the UNMODIFIED reference (oracle/ref_shim.py): a custom-SPC-like
rectangular MLA, RGB exposure + white image, patch size 9, LfpAligner (de-vignetting + local resampling) ->
LfpExtractor -> LfpRefocuser with ran_refo = [-1, 1].  Build container only:

    python oracle/gen_golden_config1.py        # writes tests/golden/config1_chain.npz

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402
from oracle.gen_golden_devignetter import white_image  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shim.load_reference()
    from plenopticam.lfp_aligner import LfpAligner
    from plenopticam.lfp_extractor import LfpExtractor
    from plenopticam.lfp_refocuser import LfpRefocuser
    M, pitch, LY, LX = 9, 11.0, 34, 37
    cent = onp.synth_centroids(LY, LX, pitch, 'rec', jitter=0.2, origin=(8.0, 8.0), seed=61)
    H, W = int(cent[:, 0].max() + 9), int(cent[:, 1].max() + 9)
    lfp = onp.synth_raw(H, W, 3, seed=62, dtype=np.uint16)
    wht = white_image(H, W, cent, pitch, 3, 63, np.uint16)
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg, sta = ref_shim.make_cfg_sta(tmp, cent, 'rec', M, ran_refo=(-1, 1))
        cfg.lfpimg = {}
        for k, v in ((cfg.opt_vign, True), (cfg.opt_rota, False), (cfg.opt_view, True), (cfg.opt_refo, True),
                     (cfg.opt_colo, False), (cfg.opt_cont, False), (cfg.opt_awb_, False), (cfg.opt_sat_, False),
                     (cfg.opt_lier, False), (cfg.opt_arti, False), (cfg.opt_dpth, False), (cfg.opt_pflu, False)):
            cfg.params[k] = v
        al = LfpAligner(lfp_img=lfp, cfg=cfg, sta=sta, wht_img=wht)
        assert al.main() is True
        aligned = np.asarray(al.lfp_img)
        ex = LfpExtractor(aligned, cfg=cfg, sta=sta)
        assert ex.main() is True
        vp = np.asarray(ex.vp_img_linear)
        rf = LfpRefocuser(vp, cfg=cfg, sta=sta)
        assert rf.main() is True
        stack = np.asarray(rf.refo_stack)
    print("aligned", aligned.shape, aligned.dtype, "vp", vp.shape, vp.dtype, "stack", stack.shape, stack.dtype,
          float(stack.min()), float(stack.max()))
    # the oracle restatement chained the same way
    dv, _, _ = onp.devignette(lfp, wht, float(M))
    o_al = onp.uint16_norm(onp.resample(dv, cent, M, 'rec'))
    o_vp = onp.compose_viewpoints(o_al, M)
    o_st = onp.refocuser_postprocess(onp.refocus_int(o_vp, (-1, 1), M))
    print("oracle == reference: aligned", np.array_equal(o_al, aligned), "vp", np.array_equal(o_vp, vp),
          "stack max abs diff", float(np.abs(o_st.astype(np.float64) - stack).max()))
    np.savez_compressed(os.path.join(GOLDEN, "config1_chain.npz"), lfp=lfp, wht=wht, centroids=cent, M=M,
                        ran_refo=(-1, 1), aligned=aligned, vp=vp, stack=stack)


if __name__ == "__main__":
    main()
