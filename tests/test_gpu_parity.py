"""Parity of the CUDA path (through the C-ABI in libpcb200.so) against the numpy oracle and against the
golden vectors produced by the unmodified reference.  All tests need a GPU.

Tolerances (BASELINE.json north_star):
  * bit-exact : viewpoint gather / inverse / crop, centroid indexing, integer shift-and-sum of uint16 data
                (compared after the float32 cast the reference's Normalizer applies next)
  * <= 1e-4 max-abs on [0,1]-normalised data for the interpolated paths; the tests use tighter figures
    where float32 allows it (stated per test).
"""
import numpy as np
import pytest

from conftest import golden, golden_names, make_cfg
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def _norm01(x, ref):
    lo, hi = float(ref.min()), float(ref.max())
    return (np.asarray(x, dtype=np.float64) - lo) / (hi - lo if hi > lo else 1.0)


# ------------------------------------------------------------------------------------------------------
# (a) resampling
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("resample_"))
def test_resample_f32_vs_reference_golden(cuda, name):
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    g = golden(name)
    M = int(g["M_used"])
    tabs = ops.LensTables(build_tables(g["centroids"], g["raw"].shape, M, str(g["pat_type"])))
    raw = ops.to_device(g["raw"])
    out, mm = ops.resample_f32(raw, tabs, flip=bool(g["flip"]), want_minmax=True)
    got = ops.to_host(out)
    ref = g["aligned_f64"]
    assert got.shape == ref.shape
    err = np.abs(_norm01(got, ref) - _norm01(ref, ref)).max()
    assert err <= 2e-6, err                                   # float32 bilinear+lerp; bar is 1e-4
    mn, mx = ops.minmax_values(mm)
    assert mn == got.min() and mx == got.max()                # fused reduction is exact


@pytest.mark.parametrize("name", golden_names("resample_"))
def test_resampler_class_u16_vs_reference_golden(cuda, name, tmp_path):
    from plenopticam_b200.lfp_aligner import LfpResampler
    g = golden(name)
    cfg, sta = make_cfg(g["centroids"], str(g["pat_type"]), int(g["M"]), tmpdir=tmp_path)
    obj = LfpResampler(lfp_img=g["raw"], cfg=cfg, sta=sta, flip=bool(g["flip"]))
    assert obj.main() is True
    got = obj.lfp_img_align
    assert got.dtype == np.uint16 and got.shape == g["aligned_u16"].shape
    assert int(cfg.params[cfg.ptc_leng]) == int(g["M_used"])
    diff = np.abs(got.astype(np.int64) - g["aligned_u16"].astype(np.int64))
    assert diff.max() <= 1, diff.max()                        # 1 LSB = 1.5e-5 of full scale (bar 1e-4)
    assert (diff != 0).mean() < 0.02
    # checkpoint side effect of the reference (lfp_resampler.py:64-68)
    import os, pickle
    with open(os.path.join(cfg.exp_path, 'lfp_img_align.pkl'), 'rb') as f:
        assert np.array_equal(pickle.load(f), got)


def test_resample_even_M_gives_zero_image(cuda):
    """Even M: the reference's window is (M+3)^2 != (M+2)^2 -> every patch is zero (lfp_local_resampler.py:50)."""
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    cent = onp.synth_centroids(5, 6, 9.0, 'rec', jitter=0.1, seed=3)
    raw = onp.synth_raw(60, 70, 3, seed=3)
    tabs = ops.LensTables(build_tables(cent, raw.shape, 6, 'rec'))
    out, _ = ops.resample_u16(ops.to_device(raw), tabs)
    assert not ops.to_host(out).any()
    assert not onp.resample(raw, cent, 6, 'rec').any()


@pytest.mark.parametrize("pat,dtype", [("hex", np.uint16), ("rec", np.float32), ("hex", np.uint8)])
def test_resample_medium_vs_oracle(cuda, pat, dtype):
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    LY, LX, pitch, M = 40, 53, 14.285, 15
    cent = onp.synth_centroids(LY, LX, pitch, pat, hex_odd=True, jitter=0.15, rot_deg=0.05, origin=(9.5, 9.5), seed=5)
    H, W = int(cent[:, 0].max() + 10), int(cent[:, 1].max() + 10)
    raw = onp.synth_raw(H, W, 3, seed=5, dtype=np.uint16 if dtype != np.float32 else np.float32)
    if dtype == np.uint8:
        raw = (raw >> 8).astype(np.uint8)
    ref = onp.resample(raw, cent, M, pat)
    tabs = ops.LensTables(build_tables(cent, raw.shape, M, pat))
    dev = ops.to_device(raw)
    assert str(dev.dtype).endswith(np.dtype(dtype).name)
    got = ops.to_host(ops.resample_f32(dev, tabs))
    assert np.abs(_norm01(got, ref) - _norm01(ref, ref)).max() <= 2e-6
    q, _ = ops.resample_u16(dev, tabs)
    d = np.abs(ops.to_host(q).astype(np.int64) - onp.uint16_norm(ref).astype(np.int64))
    assert d.max() <= 1


# ------------------------------------------------------------------------------------------------------
# (a') rotation
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rotator_hex", "rotator_rec_f32"])
def test_rotator_vs_reference_golden(cuda, name):
    from plenopticam_b200.lfp_aligner import LfpRotator
    g = golden(name)
    cfg, sta = make_cfg(g["centroids"], 'hex', 7)
    obj = LfpRotator(g["raw"], g["centroids"].tolist(), cfg=cfg, sta=sta)
    assert obj.main() is True
    assert abs(obj._rad - float(g["rad"])) < 1e-12
    assert np.abs(np.asarray(obj.centroids) - g["rot_centroids"]).max() < 1e-9
    got, ref = obj.lfp_img, g["rot_img"]
    assert got.shape == ref.shape and got.dtype == np.float64
    scale = float(np.abs(g["raw"]).max())
    assert np.abs(got - ref).max() / scale <= 5e-6            # float32 spline, FIR-truncated prefilter; bar 1e-4


def test_rotate_large_angle_and_zero(cuda):
    from plenopticam_b200 import ops
    raw = onp.synth_raw(97, 131, 3, seed=9, dtype=np.float32)
    for rad in (0.3, -1.1):
        ref = onp.rotate_img(raw.astype(np.float64), rad)
        got = ops.to_host(ops.rotate(ops.to_device(raw), rad))
        assert np.abs(got - ref).max() <= 5e-6
    from plenopticam_b200.lfp_aligner import LfpRotator
    cfg, sta = make_cfg()
    obj = LfpRotator(raw, [[0, 0, 0, 0]], rad=0.0, cfg=cfg, sta=sta)
    assert obj.main() is True and np.array_equal(obj.lfp_img, raw)        # deg == 0 -> untouched (:138)


# ------------------------------------------------------------------------------------------------------
# (b) viewpoints
# ------------------------------------------------------------------------------------------------------
def test_rearranger_vs_reference_golden_bit_exact(cuda):
    from plenopticam_b200.lfp_extractor import LfpRearranger
    g = golden("rearranger_u16")
    cfg, sta = make_cfg(M=int(g["M"]))
    obj = LfpRearranger(g["aligned"], cfg=cfg, sta=sta)
    assert obj.main() is None
    assert obj.vp_img_arr.dtype == np.uint16 and np.array_equal(obj.vp_img_arr, g["vp"])
    back = LfpRearranger(vp_img_arr=g["vp"], cfg=cfg, sta=sta)
    assert back.decompose_viewpoints() is True
    assert np.array_equal(back.lfp_img_align, g["back"])


def test_cropper_vs_reference_golden_bit_exact(cuda):
    from plenopticam_b200.lfp_extractor import LfpCropper
    g = golden("cropper_u16")
    cent = onp.synth_centroids(6, 8, 9.3, 'rec', jitter=0.0, seed=22)
    cfg, sta = make_cfg(cent, 'rec', int(g["M_out"]))
    obj = LfpCropper(lfp_img_align=g["aligned"], cfg=cfg, sta=sta)
    obj.main()
    assert np.array_equal(obj.lfp_img_align, g["cropped"])
    assert int(cfg.params[cfg.ptc_leng]) == int(g["M_used"])


@pytest.mark.parametrize("dtype,C,M,S,T", [(np.uint16, 3, 15, 21, 700), (np.float32, 3, 9, 17, 33),
                                            (np.uint8, 1, 5, 8, 9), (np.uint16, 4, 7, 6, 11)])
def test_viewpoints_roundtrip_and_oracle(cuda, dtype, C, M, S, T):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(1)
    aligned = (rng.random((S * M + 3, T * M + 2, C)) * 250).astype(dtype)      # ragged border is ignored
    ref = onp.compose_viewpoints(aligned, M)
    dev = ops.to_device(aligned)
    vp = ops.viewpoints(dev, M)
    assert np.array_equal(ops.to_host(vp), ref)
    back = ops.to_host(ops.viewpoints_inverse(vp))
    assert np.array_equal(back, aligned[:S * M, :T * M])
    if M >= 7:
        fused = ops.to_host(ops.viewpoints(dev, M, M - 4))
        two_step = onp.compose_viewpoints(onp.crop_micro_images(aligned[:S * M, :T * M], M, M - 4), M - 4)
        assert np.array_equal(fused, two_step)


# ------------------------------------------------------------------------------------------------------
# (c) refocus
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["refocus_int_u16", "refocus_int_f32"])
def test_shiftandsum_vs_reference_golden(cuda, name):
    from plenopticam_b200.lfp_refocuser import LfpShiftAndSum
    g = golden(name)
    cfg, sta = make_cfg(M=int(g["M"]), ran_refo=[int(v) for v in g["ran_refo"]])
    vp_in = g["vp"].copy()
    obj = LfpShiftAndSum(vp_img_arr=vp_in, cfg=cfg, sta=sta)
    assert obj.main() is True
    got = obj.refo_stack
    assert got.dtype == np.float64 and got.shape == g["stack"].shape
    assert np.array_equal(vp_in, g["vp"])                               # caller's array untouched
    if name.endswith("u16"):
        # exact integer accumulation: equals the reference after its next step (float32 cast in Normalizer)
        assert np.array_equal(got.astype(np.float32), g["stack"].astype(np.float32))
    else:
        assert np.abs(got - g["stack"]).max() <= 2e-6 * np.abs(g["stack"]).max()


def test_refocus_int_large_shifts_and_all_views(cuda):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(3)
    M, S, T = 9, 23, 31
    vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
    a_list = [-40, -7, -1, 0, 1, 5, 33]
    ref = np.asarray([onp.refocus_int_slice(onp._apply_aperture(vp, M) / M, a, M) for a in a_list])
    got = ops.to_host(ops.refocus_int(ops.to_device(vp), a_list))
    assert np.array_equal(got, ref.astype(np.float32))
    # no aperture: every view contributes
    none = np.zeros((M, M), dtype=bool)
    ref_all = np.zeros((S, T, 3))
    c = M // 2
    for j in range(M):
        for i in range(M):
            yi = np.clip(np.arange(S) + 2 * (c - j), 0, S - 1)
            xi = np.clip(np.arange(T) + 2 * (c - i), 0, T - 1)
            ref_all += vp[j, i][yi][:, xi]
    got_all = ops.to_host(ops.refocus_int(ops.to_device(vp), [2], view_zeroed=none))[0]
    assert np.array_equal(got_all, (ref_all / M).astype(np.float32))


def test_refocus_wide_rows_use_several_column_chunks(cuda):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(4)
    M, S, T = 5, 6, 900                                                 # T*C = 2700 > 2048 elements per CTA
    vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
    ref = onp.refocus_int(vp, (-2, 3), M)
    got = ops.to_host(ops.refocus_int(ops.to_device(vp), range(-2, 3)))
    assert np.array_equal(got, ref.astype(np.float32))


def test_percentile_matches_numpy(cuda):
    from plenopticam_b200 import ops
    import torch
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((97, 113, 3)) * 1000).astype(np.float32)
    x[3, 4, 1] = 0.0
    d = ops.to_device(x)
    for q in ([0.005], [99.9], [0.0, 50.0, 100.0], [37.123]):
        got = ops.percentile(d, q).cpu().numpy()
        assert np.array_equal(got, np.percentile(x.astype(np.float64), q)), q
    lum = ops.percentile(d, [0.005, 50.0], luma=True).cpu().numpy()
    ref = np.percentile(onp.rgb2gry(x.astype(np.float64)), [0.005, 50.0])
    assert np.abs(lum - ref).max() <= 1e-6 * np.abs(ref).max() + 1e-4
    assert torch.cuda.is_available()


def test_refocuser_vs_reference_golden(cuda):
    from plenopticam_b200.lfp_refocuser import LfpRefocuser
    g = golden("refocuser_post_u16")
    cfg, sta = make_cfg(M=int(g["M"]), ran_refo=[int(v) for v in g["ran_refo"]])
    cfg.params[cfg.opt_refo] = True
    cfg.params[cfg.opt_pflu] = False
    obj = LfpRefocuser(vp_img_arr=g["vp"], cfg=cfg, sta=sta)
    assert obj.main() is True
    got = obj.refo_stack
    assert got.dtype == np.float32 and got.shape == g["stack"].shape
    assert np.abs(got.astype(np.float64) - g["stack"]).max() <= 2e-6     # [0,1] sRGB values; bar 1e-4


def test_interrupt_returns_false(cuda):
    from plenopticam_b200.lfp_aligner import LfpResampler
    from plenopticam_b200.lfp_refocuser import LfpShiftAndSum
    g = golden("resample_rec_u16")
    cfg, sta = make_cfg(g["centroids"], 'rec', 7)
    sta.interrupt = True
    assert LfpResampler(lfp_img=g["raw"], cfg=cfg, sta=sta).main() is False
    g = golden("refocus_int_u16")
    assert LfpShiftAndSum(vp_img_arr=g["vp"], cfg=cfg, sta=sta).main() is False


# ------------------------------------------------------------------------------------------------------
# row-stationary refocus kernel vs direct kernel vs oracle
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,S,T,a_list", [
    (15, 37, 61, list(range(-32, 32))),          # Illum-like angular size, shifts far larger than the image
    (7, 19, 1300, [-3, 0, 2, 5]),                # wide rows -> several x chunks, several words per thread
    (7, 5, 1100, [-3, 0, 2, 5]),                 # two pixels per thread with <= 640 threads (found by tools/stress_parity.py)
    (5, 3, 4, list(range(-40, 40))),             # tiny image, N > 64 -> two launches, shift >> T
    (9, 1, 33, [-2, -1, 0, 1, 2]),               # single row: top and bottom edge are the same row
    (3, 40, 7, [100, -100, 1]),                  # |a| huge
])
def test_refocus_rows_equals_direct_and_oracle(cuda, M, S, T, a_list):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(M * 1000 + S)
    vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
    vp[:, :, :, :, 0] |= 0xff00                    # exercise the carry between the packed halves
    dev = ops.to_device(vp)
    mm = ops.new_minmax()
    fast = ops.to_host(ops.refocus_int(dev, a_list, mm=mm))
    direct = ops.to_host(ops.refocus_int_direct(dev, a_list))
    assert np.array_equal(fast, direct)
    vp64 = onp._apply_aperture(vp, M) / M
    for k in sorted(set([0, len(a_list) // 2, len(a_list) - 1])):
        ref = onp.refocus_int_slice(vp64, a_list[k], M).astype(np.float32)
        assert np.array_equal(fast[k], ref), (k, a_list[k])
    mn, mx = ops.minmax_values(mm)
    assert mn == fast.min() and mx == fast.max()


def test_refocus_rows_custom_mask_and_mono_fallback(cuda):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(11)
    M, S, T = 7, 9, 14
    vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
    mask = rng.random((M, M)) < 0.4
    mask[2, :] = True                               # a fully masked view row
    dev = ops.to_device(vp)
    a_list = [-4, 1, 3]
    fast = ops.to_host(ops.refocus_int(dev, a_list, view_zeroed=mask))
    direct = ops.to_host(ops.refocus_int_direct(dev, a_list, view_zeroed=mask))
    assert np.array_equal(fast, direct)
    # float32 light field -> direct kernel through the same entry point
    vpf = rng.random((M, M, S, T, 3), dtype=np.float32)
    got = ops.to_host(ops.refocus_int(ops.to_device(vpf), a_list))
    ref = np.asarray([onp.refocus_int_slice(onp._apply_aperture(vpf, M) / M, a, M) for a in a_list])
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("kernel", ["", "px", "rows", "direct"])
@pytest.mark.parametrize("mask_kind", ["circle", "contiguous_asymmetric"])
def test_refocus_kernel_variants_agree(cuda, kernel, mask_kind, monkeypatch):
    """Persistent packed-pixel (default), chunked packed-pixel, element-pair rows and direct kernels are all
    exact; a contiguous but off-centre run of views takes the packed-pixel kernels without centre-view reuse."""
    from plenopticam_b200 import ops
    rng = np.random.default_rng(21)
    M, S, T = 11, 29, 83
    vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
    vp[:, :, :, :, 1] |= 0xfff0                     # large values: 21-bit fields close to their limit
    if mask_kind == "circle":
        mask = ops.aperture_mask(M)
    else:
        mask = np.ones((M, M), dtype=bool)
        for j in range(M):
            lo = (3 * j) % 5
            mask[j, lo:lo + 2 + (j % 6)] = False    # one run per row, not centred; row widths 2..7
        mask[4, :] = True                           # a fully masked view row
    a_list = [-9, -2, 0, 1, 4, 17]
    if kernel:
        monkeypatch.setenv("PCB_REFOCUS_KERNEL", kernel)
    else:
        monkeypatch.delenv("PCB_REFOCUS_KERNEL", raising=False)
    mm = ops.new_minmax()
    got = ops.to_host(ops.refocus_int(ops.to_device(vp), a_list, view_zeroed=mask, mm=mm))
    c = M // 2
    vp64 = vp.astype(np.float64)
    for k, a in enumerate(a_list):
        ref = np.zeros((S, T, 3))
        for j in range(M):
            for i in range(M):
                if mask[j, i]:
                    continue
                yi = np.clip(np.arange(S) + a * (c - j), 0, S - 1)
                xi = np.clip(np.arange(T) + a * (c - i), 0, T - 1)
                ref += vp64[j, i][yi][:, xi]
        assert np.array_equal(got[k], (ref / M).astype(np.float32)), (kernel, a)
    mn, mx = ops.minmax_values(mm)
    assert mn == got.min() and mx == got.max()


def test_exposure_stream_matches_single_runs(cuda):
    """Pipelined host-buffer entry point (three streams) returns exactly what run() returns, in order."""
    import torch
    from plenopticam_b200.pipeline import LightFieldPipeline
    LY, LX, M = 20, 26, 9
    cent = onp.synth_centroids(LY, LX, 11.0, 'hex', hex_odd=True, jitter=0.1, origin=(7.5, 7.5), seed=2)
    H, W = int(cent[:, 0].max() + 9), int(cent[:, 1].max() + 9)
    pipe = LightFieldPipeline(cent, (H, W, 3), np.uint16, M, 'hex', ran_refo=(-3, 4))
    raws = [onp.synth_raw(H, W, 3, seed=s) for s in range(5)]
    singles = [pipe.run(r).copy() for r in raws]
    es = pipe.exposure_stream(depth=2)
    outs = {}
    for k, r in enumerate(raws):
        es.pinned_input(k)[...] = r
        assert es.submit() == k
        if k >= 1:
            outs[k - 1] = es.result(k - 1).copy()
    outs[len(raws) - 1] = es.result(len(raws) - 1).copy()
    with pytest.raises(IndexError):
        es.result(0)
    for k in range(len(raws)):
        assert np.array_equal(outs[k], singles[k]), k
    assert torch.cuda.is_available()


@pytest.mark.parametrize("bits", [8, 16])
@pytest.mark.gpu
def test_exporter_quantisation_and_quantised_exposure_stream(cuda, bits):
    """Normalizer(stack).uint16_norm() / .uint8_norm() on the device (lfp_extractor/lfp_exporter.py:134,195;
    misc/normalizer.py:43-51) and the exposure stream handing out the quantised stack."""
    from plenopticam_b200 import ops
    from plenopticam_b200.pipeline import LightFieldPipeline
    rng = np.random.default_rng(31)
    x = (rng.random((5, 17, 23, 3)) ** 2 * 3.0 - 0.4).astype(np.float32)
    mn, mx = x.min(), x.max()
    n = np.clip((x - mn) / (mx - mn), 0, 1)
    fn, dt, lv = (ops.uint16_norm, np.uint16, 65535) if bits == 16 else (ops.uint8_norm, np.uint8, 255)
    assert np.array_equal(ops.to_host(fn(ops.to_device(x))), np.asarray(np.round(n * lv), dtype=dt))
    LY, LX, M = 12, 16, 7
    cent = onp.synth_centroids(LY, LX, 9.0, 'rec', jitter=0.1, origin=(6.5, 6.5), seed=3)
    H, W = int(cent[:, 0].max() + 8), int(cent[:, 1].max() + 8)
    pipe = LightFieldPipeline(cent, (H, W, 3), np.uint16, M, 'rec', ran_refo=(-2, 2))
    raws = [onp.synth_raw(H, W, 3, seed=s) for s in range(3)]
    es = pipe.exposure_stream(depth=2, out_dtype=dt)
    for k, r in enumerate(raws):
        es.pinned_input(k)[...] = r
        es.submit()
        got = es.result(k).copy()
        f = pipe.run(r)
        ref = np.asarray(np.round(np.clip((f - f.min()) / (f.max() - f.min()), 0, 1) * lv), dtype=dt)
        assert got.dtype == dt and np.array_equal(got, ref), k


@pytest.mark.gpu
def test_contrast_bounds_single_launch_equals_numpy_percentiles(cuda):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(41)
    for shape, scale in (((97, 113, 3), 1000.0), ((434, 625, 3), 70000.0), ((5, 4, 3), 1.0)):
        x = (rng.random(shape) ** 2 * scale).astype(np.float32)
        x[0, 0] = 0.0
        got = ops.contrast_bounds(ops.to_device(x)).cpu().numpy()
        lum = (0.2126 * x[..., 0].astype(np.float64) + 0.7152 * x[..., 1].astype(np.float64)
               + 0.0722 * x[..., 2].astype(np.float64)).astype(np.float32)
        assert got[0] == np.percentile(lum.astype(np.float64), 0.005)
        assert got[1] == np.percentile(x.astype(np.float64), 99.9)
        two = np.array([ops.percentile(ops.to_device(x), [0.005], luma=True).cpu().numpy()[0],
                        ops.percentile(ops.to_device(x), [99.9]).cpu().numpy()[0]])
        assert np.array_equal(got, two)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["refocus_scratch_u16", "refocus_scratch_m7_u16", "refocus_scratch_f32"])
def test_shiftandsum_from_scratch_vs_reference_golden(cuda, name):
    """LfpShiftAndSum given the aligned image instead of the viewpoints (refo_from_scratch,
    lfp_shiftandsum.py:141-221): every view, zeros outside the light field."""
    from plenopticam_b200.lfp_refocuser import LfpShiftAndSum
    g = golden(name)
    cfg, sta = make_cfg(M=int(g["M"]), ran_refo=[int(v) for v in g["ran_refo"]])
    obj = LfpShiftAndSum(lfp_img_align=g["aligned"], cfg=cfg, sta=sta)
    assert obj.main() is True
    got = obj.refo_stack
    assert got.dtype == np.float64 and got.shape == g["stack"].shape
    if g["aligned"].dtype == np.uint16:
        assert np.array_equal(got, g["stack"].astype(np.float32).astype(np.float64))       # integer path: exact
    else:
        assert np.abs(got - g["stack"]).max() <= 2e-6 * np.abs(g["stack"]).max()
    cfg.params[cfg.opt_refi] = True
    with pytest.raises(NotImplementedError):
        LfpShiftAndSum(lfp_img_align=g["aligned"], cfg=cfg, sta=sta).main()
