"""GPU vs oracle AT BASELINE.json's own sizes (Illum-shaped exposure 5368 x 7728 x 3, 434 x 541 hex lenslets, M = 15,
64-slice stack) and on the configurations round 1 left untested (refinement at M = 15).

The oracle is run in bands / single slices so that it finishes in seconds and a bounded amount of host memory:
  * resampling: EVERY lens row of the float32 aligned image, in bands of 32 lens rows (tolerance 2e-6 of the range; the
    north star's bar is 1e-4), and the uint16 image against the oracle's Normalizer arithmetic fed with the device's
    own min / max (<= 1 LSB, < 2 % of the samples);
  * integer stack: two whole slices (a = -32 and a = 17) bit-exact against onp.refocus_int_slice;
  * refinement: M = 15, shifts up to +-30 up-sampled pixels, against onp.refocus_refi (5e-6 of the range)."""
import numpy as np
import pytest

from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def illum(cuda):
    import torch
    import bench
    from plenopticam_b200.misc import synth
    from plenopticam_b200.pipeline import LightFieldPipeline
    w = synth.illum_workload(64)
    cent = synth.illum_centroids(w)
    pipe = LightFieldPipeline(cent, (w["H"], w["W"], w["C"]), np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"])
    raw = bench.synth_raw_device(torch, w["H"], w["W"], w["C"], seed=123, device=pipe.device)
    return w, cent, pipe, raw


def _oracle_band(raw_host, cent, M, hex_odd, lo, hi):
    """Oracle resampling of lens rows lo..hi-1 (lo even keeps the row parity of the hex stretch)."""
    assert lo % 2 == 0
    sel = (cent[:, 2] >= lo) & (cent[:, 2] < hi)
    sub = cent[sel].copy()
    sub[:, 2] -= lo
    return onp.resample_hex(raw_host, sub, M, hex_odd=hex_odd)


def test_resample_every_lens_row_vs_oracle(illum):
    from plenopticam_b200 import ops
    w, cent, pipe, raw = illum
    M = w["M"]
    raw_host = ops.to_host(raw)
    hex_odd = onp.get_hex_direction(cent)
    al32, mm = ops.resample_f32(raw, pipe.tabs, want_minmax=True)
    al16, mm16 = ops.resample_u16(raw, pipe.tabs)
    mn, mx = ops.minmax_values(mm)
    assert ops.minmax_values(mm16) == (mn, mx)
    mn32, mx32 = np.float32(mn), np.float32(mx)
    worst, omin, omax, n_diff, n_tot = 0.0, np.inf, -np.inf, 0, 0
    BAND = 32
    for lo in range(0, w["LY"], BAND):
        hi = min(lo + BAND, w["LY"])
        ref = _oracle_band(raw_host, cent, M, hex_odd, lo, hi)                        # float64 (rows, 9375, 3)
        got = ops.to_host(al32[lo * M:hi * M])
        assert got.shape == ref.shape
        worst = max(worst, float(np.abs(got - ref).max()))
        r32 = ref.astype(np.float32)
        omin, omax = min(omin, float(r32.min())), max(omax, float(r32.max()))
        # uint16: the reference's Normalizer arithmetic (float32) with the device's bounds
        q_ref = np.asarray(np.round(onp.norm_fun(r32, mn32, mx32) * 65535), dtype=np.int64)
        q_got = ops.to_host(al16[lo * M:hi * M]).astype(np.int64)
        d = np.abs(q_got - q_ref)
        assert int(d.max()) <= 1, "uint16 aligned image differs by %d LSB in lens rows %d..%d" % (int(d.max()), lo, hi)
        n_diff += int((d != 0).sum())
        n_tot += d.size
    assert worst <= 2e-6 * 65535, "float32 aligned image: max abs error %g" % worst
    # the device's fused min / max are the float32 extremes of the oracle's image up to the same tolerance
    assert abs(omin - mn) <= 2e-6 * 65535 and abs(omax - mx) <= 2e-6 * 65535
    assert n_diff < 0.02 * n_tot, "%d of %d uint16 samples differ" % (n_diff, n_tot)


def test_two_full_size_slices_of_the_64_stack_bit_exact(illum):
    from plenopticam_b200 import ops
    w, cent, pipe, raw = illum
    M = w["M"]
    pipe.align(raw)
    vp = pipe.viewpoints()
    a_list = list(range(*w["ran_refo"]))
    stack = ops.refocus_int(vp, a_list, plan=pipe.refocus_plan)
    vp64 = onp._apply_aperture(ops.to_host(vp), M)
    vp64 /= M
    for a in (-32, 17):
        ref = onp.refocus_int_slice(vp64, a, M).astype(np.float32)
        got = ops.to_host(stack[a_list.index(a)])
        assert np.array_equal(got, ref), "slice a = %d: %d samples differ" % (a, int((got != ref).sum()))
    # and through the fused epilogue: slice 0 gives the bounds, the two slices follow the oracle's colour management
    fused = pipe.refocus_post()
    lin = np.stack([onp.refocus_int_slice(vp64, a, M) for a in (a_list[0], 17)])
    ref_post = onp.srgb_conv(onp.auto_hist_align(lin, lin[0]))
    got_post = np.stack([ops.to_host(fused[0]), ops.to_host(fused[a_list.index(17)])])
    assert np.abs(got_post.astype(np.float64) - ref_post).max() <= 1e-4


def test_refinement_at_m15_vs_oracle(cuda):
    """BASELINE configs[3] in miniature with the Illum angular size: M = 15, ran_refo = [-2, 2] -> shifts a in [-30, 30)
    up-sampled pixels (per-view offsets up to +-14 coarse pixels).  The oracle evaluates six of the sixty slices."""
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
    M, S, T = 15, 30, 40
    rng = np.random.default_rng(15)
    vp = (rng.random((M, M, S, T, 3)) ** 2 * 65535).astype(np.uint16)
    ran = (-2, 2)
    a_all = list(range(M * ran[0], M * ran[1]))
    probe = [-30, -17, -1, 0, 8, 29]
    ref = onp.refocus_refi(vp, ran, M, a_list=probe)
    mm = ops.new_minmax()
    got = ops.to_host(refocus_refi(ops.to_device(vp), a_all, ops.aperture_mask(M), mm=mm))
    assert got.shape == (60, S, T, 3)
    sel = got[[a_all.index(a) for a in probe]]
    assert np.abs(sel - ref).max() <= 5e-6 * np.abs(ref).max()
    mn, mx = ops.minmax_values(mm)
    assert mn == got.min() and mx == got.max()


def test_shiftandsum_refi_vs_reference_golden_negative_range(cuda):
    """The unmodified reference's own output for M = 7, ran_refo = [-1, 1] with opt_refi (oracle/gen_golden.py)."""
    from conftest import golden, make_cfg
    from plenopticam_b200.lfp_refocuser import LfpShiftAndSum
    g = golden("refocus_refi_m7_neg_u16")
    cfg, sta = make_cfg(M=int(g["M"]), ran_refo=[int(v) for v in g["ran_refo"]], refi=True)
    obj = LfpShiftAndSum(vp_img_arr=g["vp"], cfg=cfg, sta=sta)
    assert obj.main() is True
    got = obj.refo_stack
    assert got.shape == g["stack"].shape
    assert np.abs(got - g["stack"]).max() <= 5e-6 * np.abs(g["stack"]).max()
