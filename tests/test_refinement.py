"""Sub-pixel refocus refinement (opt_refi): host operator construction (CPU) and the device passes (GPU)
against the golden vectors of the unmodified reference and against the oracle."""
import numpy as np
import pytest

from conftest import golden, make_cfg
from oracle import oracle_np as onp


def test_spline_matrix_equals_fitpack_interpolating_spline():
    from plenopticam_b200.lfp_refocuser import refinement as rf
    for n_in, n_out in [(9, 45), (45, 9), (10, 50), (4, 20), (33, 7)]:
        W = rf.spline_matrix(n_in, n_out).toarray()
        assert np.abs(W - onp.spline_resize_matrix(n_in, n_out)).max() < 1e-13
    with pytest.raises(NotImplementedError):
        rf.spline_matrix(3, 9)


def test_band_operators_reproduce_reference_refinement():
    """R_a = (1/M) sum A V B^T with the truncated float32 bands == the reference's upsample/shift/sum/downsample."""
    from plenopticam_b200.lfp_refocuser import refinement as rf
    g = golden("refocus_refi_u16")
    vp, M = g["vp"], int(g["M"])
    ran = [int(v) for v in g["ran_refo"]]
    a_list = list(range(M * ran[0], M * ran[1]))
    S, T = vp.shape[2:4]
    Ay, Bx = rf.build_operators(S, T, M, a_list)
    mask, c = onp.aperture_mask(M), M // 2
    for n in (0, len(a_list) // 2, len(a_list) - 1):
        a = a_list[n]
        acc = np.zeros((S, T, 3))
        for j in range(M):
            for i in range(M):
                if not mask[j, i]:
                    acc += Ay.apply(a * (c - j), Bx.apply(a * (c - i), vp[j, i].astype(np.float64), axis=1), axis=0)
        assert np.abs(acc / M - g["stack"][n]).max() <= 1e-6 * np.abs(g["stack"]).max()
    # device layouts == plain band (applied to B-spline coefficients)
    rng = np.random.default_rng(0)
    vals, first, Kg = Ay.grouped()                              # vertical pass: groups of 4 outputs share a window
    f = rng.random(S)
    for s in Ay.shifts[::7]:
        k = Ay.index[s]
        out = np.zeros(vals.shape[1] * 4)
        for gi in range(vals.shape[1]):
            for m in range(Kg):
                out[4 * gi:4 * gi + 4] += vals[k, gi, m] * f[min(first[k, gi] + m, S - 1)]
        assert np.abs(out[:S] - Ay.apply_coef(s, f)).max() < 1e-6
    tv, tf, lo, hi = Bx.tap_major()                             # horizontal pass: one thread per output, tap-major weights
    f = rng.random(T)
    for s in Bx.shifts[::7]:
        k = Bx.index[s]
        out = np.zeros(T)
        for x in range(T):
            assert lo[k] <= tf[k, x] - x and tf[k, x] - x + Bx.K - 1 <= hi[k]
            for m in range(Bx.K):
                out[x] += tv[k, m, x] * f[tf[k, x] + m]
        assert np.abs(out - Bx.apply_coef(s, f)).max() < 1e-6
    # the factorisation itself: D_s . P == Wdn . Shift_s . Wup
    nf = T * M
    full = (rf.spline_matrix(nf, T) @ rf.spline_matrix(T, nf)[np.clip(np.arange(nf) + Bx.shifts[1], 0, nf - 1)]).toarray()
    assert np.abs(Bx.apply(Bx.shifts[1], np.eye(T)) - full).max() < 1e-6
    assert Bx.K <= 8 and Ay.K <= 8


def test_refinement_band_is_narrow_at_illum_size():
    """At Lytro-Illum size the per-slice band is 5 taps (the prefilter carries the long tails)."""
    from plenopticam_b200.lfp_refocuser import refinement as rf
    ops_x = rf.AxisOperators(625, 15, [0, 7, -105, 224])
    assert ops_x.K <= 5 and 20 <= ops_x.Kp <= 32
    f = np.random.default_rng(1).random(625)
    nf = 625 * 15
    for s in ops_x.shifts:
        full = rf.spline_matrix(nf, 625) @ (rf.spline_matrix(625, nf) @ f)[np.clip(np.arange(nf) + s, 0, nf - 1)]
        assert np.abs(ops_x.apply(s, f) - full).max() < 1e-6


@pytest.mark.gpu
def test_shiftandsum_refi_vs_reference_golden(cuda):
    from plenopticam_b200.lfp_refocuser import LfpShiftAndSum
    g = golden("refocus_refi_u16")
    cfg, sta = make_cfg(M=int(g["M"]), ran_refo=[int(v) for v in g["ran_refo"]], refi=True)
    obj = LfpShiftAndSum(vp_img_arr=g["vp"], cfg=cfg, sta=sta)
    assert obj.main() is True
    got = obj.refo_stack
    assert got.shape == g["stack"].shape and got.dtype == np.float64
    # float32 banded evaluation; bar: 1e-4 of the value range
    assert np.abs(got - g["stack"]).max() <= 5e-6 * np.abs(g["stack"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_refi_medium_vs_oracle(cuda, dtype):
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
    rng = np.random.default_rng(12)
    M, S, T = 7, 21, 30
    vp = (rng.random((M, M, S, T, 3)) * (65535 if dtype == np.uint16 else 1)).astype(dtype)
    ran = (-1, 2)
    ref = onp.refocus_refi(vp, ran, M)
    mm = ops.new_minmax()
    got = ops.to_host(refocus_refi(ops.to_device(vp), range(M * ran[0], M * ran[1]), ops.aperture_mask(M), mm=mm))
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 5e-6 * np.abs(ref).max()
    mn, mx = ops.minmax_values(mm)
    assert mn == got.min() and mx == got.max()
