"""Golden vectors of the UNMODIFIED reference LfpContrast.main (extractor colour management with its options off;
through oracle/ref_shim.py).  Build container only:  python oracle/gen_golden_contrast.py
writes tests/golden/vp_contrast.npz.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shim.load_reference()
    from plenopticam.lfp_extractor.lfp_contrast import LfpContrast
    rng = np.random.default_rng(91)
    out = {}
    for name, M, S, T, zero_lo in (("plain", 5, 14, 18, False), ("lo_zero", 3, 12, 16, True)):
        vp = (rng.random((M, M, S, T, 3)) ** 2 * 60000 + 300).astype(np.uint16)
        if zero_lo:
            vp[M // 2, M // 2, :2, 0, :] = 0          # the two smallest luma samples of the central view are 0 -> lo == 0
        with tempfile.TemporaryDirectory() as tmp:
            cfg, sta = ref_shim.make_cfg_sta(tmp, onp.synth_centroids(3, 3, 9.0, 'rec', jitter=0, seed=0), 'rec', M)
            for k in ('opt_cont', 'opt_awb_', 'opt_sat_'):
                cfg.params[getattr(cfg, k)] = False
            obj = LfpContrast(vp_img_arr=vp.copy(), cfg=cfg, sta=sta)
            assert obj.main() is True
            ref = np.asarray(obj.vp_img_arr)
        mine = onp.srgb_conv(onp.auto_hist_align(vp, vp[M // 2, M // 2].astype('float64')))
        print(name, ref.dtype, ref.shape, "oracle == reference:", np.array_equal(mine, ref), float(np.abs(mine - ref).max()),
              "lo, hi =", onp.hist_percentiles(vp[M // 2, M // 2].astype('float64')))
        out[name + "_vp"] = vp
        out[name + "_out"] = ref
    np.savez_compressed(os.path.join(GOLDEN, "vp_contrast.npz"), **out)


if __name__ == "__main__":
    main()
