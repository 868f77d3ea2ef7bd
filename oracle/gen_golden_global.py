"""Golden vectors of the UNMODIFIED reference LfpResampler with smp_meth='global' (through oracle/ref_shim.py).
skimage is absent, so `transform.warp` is served by the oracle's restatement (oracle_np.warp_projective): everything
but that one call is the reference's own code.  Build container only:

    python oracle/gen_golden_global.py        # writes tests/golden/global_*.npz

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

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shim.load_reference()
    from plenopticam.lfp_aligner.lfp_resampler import LfpResampler
    for name, pat, LY, LX, ho, pitch in (("global_rec_u16", 'rec', 9, 11, True, 11.3),
                                         ("global_hex_odd_u16", 'hex', 10, 12, True, 11.3),
                                         ("global_hex_even_u16", 'hex', 11, 10, False, 12.6)):
        cent = onp.synth_centroids(LY, LX, pitch, pat, hex_odd=ho, jitter=0.1, rot_deg=0.05, origin=(9., 9.), seed=71)
        H, W = int(cent[:, 0].max() + 10), int(cent[:, 1].max() + 10)
        raw = onp.synth_raw(H, W, 3, seed=72)
        with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cfg, sta = ref_shim.make_cfg_sta(tmp, cent, pat, 9, smp_meth='global')
            obj = LfpResampler(lfp_img=raw, cfg=cfg, sta=sta)
            obj.global_resampling()
            f64 = np.asarray(obj._lfp_img_align).copy()
            pitch_out = np.asarray(obj._limg_pitch).copy()
            obj2 = LfpResampler(lfp_img=raw, cfg=cfg, sta=sta)
            assert obj2.main() is True
            u16 = np.asarray(obj2.lfp_img_align)
        mine, pit = onp.resample_global(raw, cent, pat)
        print(name, f64.shape, u16.dtype, pitch_out, "oracle == reference:", float(np.abs(mine - f64).max()),
              np.array_equal(onp.uint16_norm(mine), u16))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), raw=raw, centroids=cent, pat=pat, aligned_f64=f64,
                            aligned_u16=u16, pitch=pitch_out)


if __name__ == "__main__":
    main()
