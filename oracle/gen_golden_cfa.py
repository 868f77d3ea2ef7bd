"""Golden vectors of the UNMODIFIED reference CfaProcessor.safe_bayer_awb (through oracle/ref_shim.py).
Build container only:  python oracle/gen_golden_cfa.py   # writes tests/golden/cfa_awb.npz
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = (   # name, (H, W), pattern, awb, exp, white image?
    ("grbg_wht_exp", (36, 52), "GRBG", [1.31, 1.0, 1.62, 1.0], -0.5, True),
    ("bggr_nowht", (28, 40), "BGGR", [1.9, 1.05, 1.4, 1.0], None, False),
    ("grbg_wht0_exp", (20, 24), "GRBG", [2.2, 1.0, 1.7, 1.02], 0.75, True),
    ("bggr_wht0", (30, 44), "BGGR", [1.7, 1.0, 2.1, 1.0], None, True),
)


def main():
    ref_shim.load_reference()
    from plenopticam.lfp_aligner.cfa_processor import CfaProcessor
    rng = np.random.default_rng(77)
    out = {}
    for name, (H, W), pat, awb, exp, has_wht in CASES:
        bay = (rng.random((H, W)) * 1.15).astype(np.float32)        # some samples beyond the saturation threshold
        bay[rng.random((H, W)) < 0.02] = 0.0
        wht = None
        if has_wht:
            wht = (0.45 + 0.7 * rng.random((H, W))).astype(np.float32)
            if "wht0" in name:
                wht[rng.random((H, W)) < 0.05] = 0.0                  # dead white pixels: reference threshold -> 0
        with tempfile.TemporaryDirectory() as tmp:
            cfg, sta = ref_shim.make_cfg_sta(tmp, onp.synth_centroids(3, 3, 9.0, 'rec', jitter=0, seed=0), 'rec', 7)
            cfg.lfpimg = {'bit': 10, 'bay': pat, 'awb': list(awb)}
            if exp is not None:
                cfg.lfpimg['exp'] = exp
            obj = CfaProcessor(bay_img=bay.copy(), wht_img=None if wht is None else wht.copy(), cfg=cfg, sta=sta)
            obj.safe_bayer_awb()
            ref = np.asarray(obj._bay_img)
        mine = onp.cfa_safe_bayer_awb(bay, wht, awb, pat, exp)
        print(name, ref.dtype, ref.shape, "oracle == reference:", np.array_equal(mine, ref),
              float(np.abs(mine - ref).max()))
        out[name + "_bay"] = bay
        if wht is not None:
            out[name + "_wht"] = wht
        out[name + "_awb"] = np.array(awb)
        out[name + "_exp"] = np.array([np.nan if exp is None else exp])
        out[name + "_out"] = ref
    np.savez_compressed(os.path.join(GOLDEN, "cfa_awb.npz"), **out)


if __name__ == "__main__":
    main()
