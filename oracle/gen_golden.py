"""Generate golden input/output vectors by running the UNMODIFIED Python reference (through
`oracle/ref_shim.py`) on small seeded synthetic cases.  Run in the build container only:

    python oracle/gen_golden.py            # writes tests/golden/*.npz

TEST INFRASTRUCTURE ONLY.  The reference cannot travel to the GPU box, so the vectors are committed.
Each .npz stores the inputs and what the reference classes returned:
  LfpResampler.main()      -> lfp_img_align (uint16) and the float64 array before quantisation
  LfpRotator.main()        -> lfp_img, centroids
  LfpRearranger.main()     -> vp_img_arr
  LfpCropper.main()        -> lfp_img_align
  LfpShiftAndSum.main()    -> refo_stack (float64), integer and opt_refi
  LfpRefocuser.shift_and_sum post-processing -> float32 sRGB stack
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


def _raw_for(cent, pitch, C=3, seed=0, dtype=np.uint16, margin=None):
    H = int(np.ceil(cent[:, 0].max() + (pitch if margin is None else margin)))
    W = int(np.ceil(cent[:, 1].max() + (pitch if margin is None else margin)))
    return onp.synth_raw(H, W, C, seed=seed, dtype=dtype)


def run_resampler(raw, cent, M, pat_type, flip=False):
    ref_shim.load_reference()
    from plenopticam.lfp_aligner import LfpResampler
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, cent, pat_type, M)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            obj = LfpResampler(lfp_img=raw, cfg=cfg, sta=sta, flip=flip)
            # float64 result before the uint16 quantisation of _write_lfp_align
            obj.local_resampling()
            f64 = obj.lfp_img_align
            obj2 = LfpResampler(lfp_img=raw, cfg=cfg, sta=sta, flip=flip)
            ret = obj2.main()
        assert ret is True
        return f64, obj2.lfp_img_align, int(cfg.params[cfg.ptc_leng])


def run_rotator(raw, cent, rad=None):
    ref_shim.load_reference()
    from plenopticam.lfp_aligner.lfp_rotator import LfpRotator
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, cent, 'hex', 7)
        obj = LfpRotator(raw.astype('float'), np.array(cent).tolist(), rad=rad, cfg=cfg, sta=sta)
        assert obj.main() is True
        return obj.lfp_img, np.asarray(obj.centroids), float(obj._rad)


def run_rearranger(aligned, M):
    ref_shim.load_reference()
    from plenopticam.lfp_extractor.lfp_rearranger import LfpRearranger
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, [[0, 0, 0, 0]], 'rec', M)
        obj = LfpRearranger(aligned, cfg=cfg, sta=sta)
        obj.main()
        vp = obj.vp_img_arr
        back = LfpRearranger(vp_img_arr=vp, cfg=cfg, sta=sta)
        back._dtype = vp.dtype
        back.decompose_viewpoints()
        return vp, back._lfp_img_align


def run_cropper(aligned, cent, pat_type, M_out):
    ref_shim.load_reference()
    from plenopticam.lfp_extractor.lfp_cropper import LfpCropper
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, cent, pat_type, M_out)
        obj = LfpCropper(lfp_img_align=aligned, cfg=cfg, sta=sta)
        obj.main()
        return obj.lfp_img_align, int(cfg.params[cfg.ptc_leng])


def run_shiftandsum(vp, M, ran_refo, refi=False):
    ref_shim.load_reference()
    from plenopticam.lfp_refocuser.lfp_shiftandsum import LfpShiftAndSum
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, [[0, 0, 0, 0]], 'rec', M, ran_refo=ran_refo, opt_refi=refi)
        obj = LfpShiftAndSum(vp_img_arr=vp, cfg=cfg, sta=sta)
        assert obj.main() is True
        return obj.refo_stack


def run_refocuser(vp, M, ran_refo):
    """LfpRefocuser.main() including auto_hist_align + sRGB (exports are stubbed by the shim)."""
    ref_shim.load_reference()
    from plenopticam.lfp_refocuser import LfpRefocuser
    with tempfile.TemporaryDirectory() as tmp:
        cfg, sta = ref_shim.make_cfg_sta(tmp, [[0, 0, 0, 0]], 'rec', M, ran_refo=ran_refo)
        cfg.params[cfg.opt_refo] = True
        cfg.params[cfg.opt_pflu] = False
        obj = LfpRefocuser(vp_img_arr=vp, cfg=cfg, sta=sta)
        assert obj.main() is True
        return np.asarray(obj.refo_stack)


def main():
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- resampler: rec / hex (both row parities) / flip / float raw / border lenslets / 1 channel
    cases = [
        dict(name="resample_rec_u16", pat='rec', LY=6, LX=8, pitch=9.3, M=7, seed=1, dtype=np.uint16),
        dict(name="resample_hex_odd_u16", pat='hex', LY=7, LX=9, pitch=9.6, M=7, seed=2, dtype=np.uint16, hex_odd=True),
        dict(name="resample_hex_even_f32", pat='hex', LY=6, LX=10, pitch=11.2, M=9, seed=3, dtype=np.float32, hex_odd=False),
        dict(name="resample_rec_flip_f32", pat='rec', LY=5, LX=5, pitch=7.4, M=5, seed=4, dtype=np.float32, flip=True),
        dict(name="resample_hex_border_u16", pat='hex', LY=6, LX=7, pitch=9.0, M=9, seed=5, dtype=np.uint16, hex_odd=True,
             origin=(4.2, 4.4), margin=3),
        dict(name="resample_rec_mono_u16", pat='rec', LY=5, LX=6, pitch=8.2, M=7, seed=6, dtype=np.uint16, C=1),
        dict(name="resample_hex_rot_u16", pat='hex', LY=8, LX=9, pitch=10.1, M=9, seed=7, dtype=np.uint16, hex_odd=True,
             rot=0.4),
    ]
    for cs in cases:
        cent = onp.synth_centroids(cs['LY'], cs['LX'], cs['pitch'], cs['pat'], hex_odd=cs.get('hex_odd', True),
                                   jitter=0.2, rot_deg=cs.get('rot', 0.0), origin=cs.get('origin'), seed=cs['seed'])
        raw = _raw_for(cent, cs['pitch'], C=cs.get('C', 3), seed=cs['seed'], dtype=cs['dtype'], margin=cs.get('margin'))
        # (a 2-D mono image crashes the reference at lfp_local_resampler.py:54; (H,W,1) works)
        f64, u16, M_used = run_resampler(raw, cent, cs['M'], cs['pat'], flip=cs.get('flip', False))
        np.savez_compressed(os.path.join(GOLDEN, cs['name'] + ".npz"), raw=raw, centroids=cent, M=cs['M'],
                            M_used=M_used, pat_type=cs['pat'], flip=cs.get('flip', False),
                            aligned_f64=f64, aligned_u16=u16)
        print(cs['name'], raw.shape, f64.shape, u16.dtype, "M_used", M_used)

    # ---- rotator
    cent = onp.synth_centroids(9, 9, 10.0, 'hex', hex_odd=True, jitter=0.1, rot_deg=0.6, seed=11)
    raw = _raw_for(cent, 10.0, seed=11, dtype=np.uint16)
    img, cen, rad = run_rotator(raw, cent)
    np.savez_compressed(os.path.join(GOLDEN, "rotator_hex.npz"), raw=raw, centroids=cent, rot_img=img,
                        rot_centroids=cen, rad=rad)
    print("rotator", raw.shape, rad)
    cent = onp.synth_centroids(7, 8, 9.0, 'rec', jitter=0.1, rot_deg=-1.3, seed=12)
    raw = _raw_for(cent, 9.0, seed=12, dtype=np.float32)
    img, cen, rad = run_rotator(raw, cent)
    np.savez_compressed(os.path.join(GOLDEN, "rotator_rec_f32.npz"), raw=raw, centroids=cent, rot_img=img,
                        rot_centroids=cen, rad=rad)
    print("rotator", raw.shape, rad)

    # ---- rearranger / cropper
    rng = np.random.default_rng(21)
    aligned = rng.integers(0, 65536, size=(6 * 7, 8 * 7, 3), dtype=np.uint16)
    vp, back = run_rearranger(aligned, 7)
    np.savez_compressed(os.path.join(GOLDEN, "rearranger_u16.npz"), aligned=aligned, M=7, vp=vp, back=back)
    print("rearranger", vp.shape, vp.dtype, np.array_equal(back, aligned))
    cent = onp.synth_centroids(6, 8, 9.3, 'rec', jitter=0.0, seed=22)
    aligned9 = rng.integers(0, 65536, size=(6 * 9, 8 * 9, 3), dtype=np.uint16)
    cropped, M_used = run_cropper(aligned9, cent, 'rec', 5)
    np.savez_compressed(os.path.join(GOLDEN, "cropper_u16.npz"), aligned=aligned9, M_in=9, M_out=5, cropped=cropped,
                        M_used=M_used)
    print("cropper", cropped.shape, M_used)

    # ---- shift and sum (integer), incl. negative a, M=5 and M=7; uint16 and float inputs
    vp_u16 = rng.integers(0, 65536, size=(7, 7, 12, 14, 3), dtype=np.uint16)
    st = run_shiftandsum(vp_u16, 7, (-3, 4))
    np.savez_compressed(os.path.join(GOLDEN, "refocus_int_u16.npz"), vp=vp_u16, M=7, ran_refo=(-3, 4), stack=st)
    print("refocus int", st.shape, st.dtype)
    vp_f32 = rng.random((5, 5, 9, 11, 3), dtype=np.float32)
    st = run_shiftandsum(vp_f32, 5, (-6, 7))
    np.savez_compressed(os.path.join(GOLDEN, "refocus_int_f32.npz"), vp=vp_f32, M=5, ran_refo=(-6, 7), stack=st)
    print("refocus int f32", st.shape)

    # ---- refinement
    vp_u16s = rng.integers(0, 65536, size=(5, 5, 9, 10, 3), dtype=np.uint16)
    st = run_shiftandsum(vp_u16s, 5, (-1, 2), refi=True)
    np.savez_compressed(os.path.join(GOLDEN, "refocus_refi_u16.npz"), vp=vp_u16s, M=5, ran_refo=(-1, 2), stack=st)
    print("refocus refi", st.shape)

    # ---- refocuser post-process (contrast + sRGB); slices >= 32 px (watermark in the stubbed exporter path)
    vp_u16p = (rng.random((5, 5, 36, 40, 3)) ** 2 * 65535).astype(np.uint16)
    st = run_refocuser(vp_u16p, 5, (-1, 2))
    np.savez_compressed(os.path.join(GOLDEN, "refocuser_post_u16.npz"), vp=vp_u16p, M=5, ran_refo=(-1, 2), stack=st)
    print("refocuser", st.shape, st.dtype, st.min(), st.max())

    # ---- refinement, round 2: M = 7 with a range that starts negative (shifts a in [-7, 7) up-sampled pixels);
    #      its own generator so that every file above stays byte-identical
    rng7 = np.random.default_rng(77)
    vp_u16r = rng7.integers(0, 65536, size=(7, 7, 11, 13, 3), dtype=np.uint16)
    st = run_shiftandsum(vp_u16r, 7, (-1, 1), refi=True)
    np.savez_compressed(os.path.join(GOLDEN, "refocus_refi_m7_neg_u16.npz"), vp=vp_u16r, M=7, ran_refo=(-1, 1), stack=st)
    print("refocus refi M=7 negative range", st.shape)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "refi_m7":          # only the round-2 addition
        _main = main

        def main():
            rng7 = np.random.default_rng(77)
            vp_u16r = rng7.integers(0, 65536, size=(7, 7, 11, 13, 3), dtype=np.uint16)
            st = run_shiftandsum(vp_u16r, 7, (-1, 1), refi=True)
            np.savez_compressed(os.path.join(GOLDEN, "refocus_refi_m7_neg_u16.npz"), vp=vp_u16r, M=7, ran_refo=(-1, 1),
                                stack=st)
            print("refocus refi M=7 negative range", st.shape)
    main()
