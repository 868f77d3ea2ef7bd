"""Golden vector of the UNMODIFIED reference LfpCropper on a light field with a NON-SQUARE micro-image pitch
(13 x 15, what smp_meth='global' yields for a hexagonal MLA).  Build container only:
    python oracle/gen_golden_cropper_yx.py     # writes tests/golden/cropper_hex_13x15_u16.npz
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


def main():
    ref_shim.load_reference()
    from plenopticam.lfp_extractor.lfp_cropper import LfpCropper
    LY, LX, py, px, M = 20, 26, 13, 15, 9
    cent = onp.synth_centroids(LY, LX, 14.0, 'hex', hex_odd=True, jitter=0.0, seed=81)
    Xs = int(round(LX * 2 / np.sqrt(3)))
    rng = np.random.default_rng(82)
    aligned = rng.integers(0, 65536, size=(LY * py, Xs * px, 3), dtype=np.uint16)
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, cent, 'hex', M)
        obj = LfpCropper(lfp_img_align=aligned, cfg=cfg, sta=sta)
        obj.main()
        out = np.asarray(obj.lfp_img_align)
        pitch = np.asarray(obj._limg_pitch)
    print(aligned.shape, "->", out.shape, "pitch", pitch, "oracle == reference:",
          np.array_equal(onp.crop_micro_images(aligned, pitch, M), out))
    np.savez_compressed(os.path.join(GOLDEN, "cropper_hex_13x15_u16.npz"), aligned=aligned, centroids=cent, M=M,
                        pitch=pitch, cropped=out)


if __name__ == "__main__":
    main()
