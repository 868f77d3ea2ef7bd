"""Golden vectors of the UNMODIFIED reference `LfpShiftAndSum.refo_from_scratch` (lfp_refocuser/lfp_shiftandsum.py:141-221),
the alternate taken when the class is given the aligned image and no viewpoints (through oracle/ref_shim.py).  Build
container only:

    python oracle/gen_golden_scratch.py        # writes tests/golden/refocus_scratch_*.npz

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


def run(aligned, M, ran_refo):
    ref_shim.load_reference()
    from plenopticam.lfp_refocuser.lfp_shiftandsum import LfpShiftAndSum
    with tempfile.TemporaryDirectory() as tmp, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg, sta = ref_shim.make_cfg_sta(tmp, [[0, 0, 0, 0]], 'rec', M, ran_refo=ran_refo)
        obj = LfpShiftAndSum(lfp_img_align=aligned, cfg=cfg, sta=sta)
        assert obj.main() is True
        return np.asarray(obj.refo_stack)


def main():
    rng = np.random.default_rng(141)
    for name, M, H, J, P, ran, dtype in (("refocus_scratch_u16", 5, 8, 9, 3, (-2, 3), np.uint16),
                                         ("refocus_scratch_m7_u16", 7, 6, 7, 3, (-1, 2), np.uint16),
                                         ("refocus_scratch_f32", 3, 7, 6, 3, (-3, 2), np.float32)):
        if np.issubdtype(dtype, np.integer):
            aligned = rng.integers(0, 65536, size=(H * M, J * M, P), dtype=dtype)
        else:
            aligned = rng.random((H * M, J * M, P), dtype=np.float32)
        stack = run(aligned, M, ran)
        o = onp.refocus_from_scratch(aligned, ran, M)
        print(name, stack.shape, stack.dtype, "oracle max rel diff", float(np.abs(o - stack).max() / np.abs(stack).max()))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), aligned=aligned, M=M, ran_refo=ran, stack=stack)


if __name__ == "__main__":
    main()
