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
    # grouped device layout == plain band
    vals, first, Kg = Bx.grouped()
    f = np.random.default_rng(0).random(T)
    for s in Bx.shifts[::7]:
        k = Bx.index[s]
        out = np.zeros(vals.shape[1] * 4)
        for gi in range(vals.shape[1]):
            for m in range(Kg):
                out[4 * gi:4 * gi + 4] += vals[k, gi, m] * f[min(first[k, gi] + m, T - 1)]
        assert np.abs(out[:T] - Bx.apply(s, f)).max() < 1e-6


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
