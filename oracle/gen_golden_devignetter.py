"""Golden vectors of the UNMODIFIED reference LfpDevignetter (through oracle/ref_shim.py).  Build container only:

    python oracle/gen_golden_devignetter.py        # writes tests/golden/devignetter_*.npz

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


from plenopticam_b200.misc.synth import white_image  # noqa: E402,F401


def run(lfp, wht, cent, M):
    ref_shim.load_reference()
    from plenopticam.lfp_aligner.lfp_devignetter import LfpDevignetter
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, cent, 'rec', M)
        obj = LfpDevignetter(lfp_img=lfp, wht_img=wht, cfg=cfg, sta=sta)
        assert obj.main() is True
        return np.asarray(obj.lfp_img), np.asarray(obj.wht_img), float(obj.noise_lev), float(np.mean(cfg.calibs[cfg.ptc_mean]))


def main():
    M, pitch = 9, 11.0
    cent = onp.synth_centroids(7, 9, pitch, 'rec', jitter=0.1, origin=(8.0, 8.0), seed=31)
    H, W = int(cent[:, 0].max() + 9), int(cent[:, 1].max() + 9)
    for name, C, ldt, wdt in [("devignetter_rgb_u16", 3, np.uint16, np.uint16),
                              ("devignetter_mono_f64", 1, np.float64, np.float64)]:
        lfp = onp.synth_raw(H, W, 3, seed=32, dtype=np.uint16)
        lfp = lfp if C == 3 else lfp[..., 0].astype(np.float64) / 7.0
        wht = white_image(H, W, cent, pitch, C, 33, wdt)
        out, wn, noise, pm = run(lfp, wht, cent, M)
        o_out, o_wn, o_noise = onp.devignette(lfp, wht, pm)
        print(name, out.shape, wn.shape, noise, "oracle == reference:", np.array_equal(o_out, out),
              np.array_equal(o_wn, wn), abs(o_noise - noise))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), lfp=lfp, wht=wht, centroids=cent, M=M,
                            pitch_mean=pm, out=out, wht_out=wn, noise_lev=noise)


if __name__ == "__main__":
    main()
