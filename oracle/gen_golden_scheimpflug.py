"""Golden vectors of the UNMODIFIED reference LfpScheimpflug.scheimpflug_from_stack (through oracle/ref_shim.py;
the image handed to misc.save_img_file is captured).  Build container only:

    python oracle/gen_golden_scheimpflug.py        # writes tests/golden/scheimpflug_f32.npz

TEST INFRASTRUCTURE ONLY.
"""
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
    import plenopticam.lfp_refocuser.lfp_scheimpflug as mod
    rng = np.random.default_rng(41)
    stack = rng.random((9, 23, 31, 3)).astype(np.float32)
    cent = onp.synth_centroids(5, 5, 9.0, 'rec', jitter=0.0, seed=1)
    captured = {}
    orig = mod.misc.save_img_file
    mod.misc.save_img_file = lambda img, file_path=None, **k: captured.__setitem__(os.path.basename(file_path), np.array(img))
    try:
        with tempfile.TemporaryDirectory() as tmp:
            cfg, sta = ref_shim.make_cfg_sta(tmp, cent, 'rec', 7, ran_refo=(0, 9))
            obj = mod.LfpScheimpflug(refo_stack=stack, cfg=cfg, sta=sta)
            assert obj.main() is True
    finally:
        mod.misc.save_img_file = orig
    out = {}
    for d in onp.PFLU_VALS:
        key = [k for k in captured if k.endswith('_' + d + '.png')][0]
        out[d.replace(' ', '_')] = captured[key]
        img, q = onp.scheimpflug_from_stack(stack, d)
        print(d, captured[key].shape, captured[key].dtype, "oracle == reference:", np.array_equal(q, captured[key]))
    np.savez_compressed(os.path.join(GOLDEN, "scheimpflug_f32.npz"), stack=stack, **out)


if __name__ == "__main__":
    main()
