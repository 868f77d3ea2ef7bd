"""Golden vectors of the `upscale_*.png` side effect of the UNMODIFIED reference's refinement path
(lfp_refocuser/lfp_shiftandsum.py:126-132: per refined slice the summed UP-SAMPLED image, histogram-aligned against itself,
sRGB-encoded, handed to LfpExporter.save_refo_slice -> misc.save_img_file -> Normalizer.uint8_norm -> imageio.imwrite).
The shim's imageio stand-in is given an `imwrite` that records what the reference would have written.  The up-sampled slices
are kept below 24 px per side so that the 16 x 16 watermark of misc/file_rw.py:143-158 cannot be placed (ValueError, caught
at :56-59) and the recorded bytes are the image alone.  Build container only:

    python oracle/gen_golden_upscale.py        # writes tests/golden/refocus_upscale_u16.npz

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
    import imageio                                   # the shim's stand-in module
    from plenopticam.lfp_refocuser.lfp_shiftandsum import LfpShiftAndSum
    written = []
    imageio.imwrite = lambda uri=None, im=None, **k: written.append((os.path.basename(uri), np.array(im)))
    rng = np.random.default_rng(126)
    M, S, T, ran = 5, 4, 4, (-1, 1)
    vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg, sta = ref_shim.make_cfg_sta(tmp, [[0, 0, 0, 0]], 'rec', M, ran_refo=ran, opt_refi=True)
        obj = LfpShiftAndSum(vp_img_arr=vp, cfg=cfg, sta=sta)
        assert obj.main() is True
        stack = np.asarray(obj.refo_stack)
    names = [w[0] for w in written]
    imgs = np.stack([w[1] for w in written])
    print("files", names, imgs.shape, imgs.dtype, "stack", stack.shape)
    # oracle: summed up-sampled image -> auto_hist_align(ref = itself) -> sRGB -> uint8_norm
    ups = onp.refocus_refi_upscaled(vp, ran, M)
    o8 = np.stack([onp.uint8_norm(onp.srgb_conv(onp.auto_hist_align(u, u))) for u in ups])
    print("oracle == reference:", np.array_equal(o8, imgs), "max diff", int(np.abs(o8.astype(int) - imgs.astype(int)).max()))
    np.savez_compressed(os.path.join(GOLDEN, "refocus_upscale_u16.npz"), vp=vp, M=M, ran_refo=ran, names=np.array(names),
                        upscaled_u8=imgs, stack=stack)


if __name__ == "__main__":
    main()
