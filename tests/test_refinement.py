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
    vals, first, Kg = Ay.grouped()                              # vertical pass: groups of GROUP outputs share a window
    f = rng.random(S)
    for s in Ay.shifts[::7]:
        k = Ay.index[s]
        out = np.zeros(vals.shape[1] * rf.GROUP)
        for gi in range(vals.shape[1]):
            assert first[k, gi] >= 0 and first[k, gi] + Kg <= S             # windows inside the image: no clamping on the device
            for m in range(Kg):
                out[rf.GROUP * gi:rf.GROUP * (gi + 1)] += vals[k, gi, m] * f[first[k, gi] + m]
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


def test_prefilter_interior_is_one_symmetric_filter():
    """Away from the border the rows of P are the cubic B-spline prefilter sqrt(3) (sqrt(3) - 2)^|k|, exactly equal as
    float32: what the register-window kernels rely on."""
    from plenopticam_b200.lfp_refocuser import refinement as rf
    ops_ = rf.AxisOperators(96, 5, [0])
    h, c, lo, hi = ops_.prefilter_toeplitz()
    assert h is not None and h.size == 29 and c == 14 and np.array_equal(h, h[::-1])
    assert lo <= 30 and hi >= 96 - 31
    z1 = np.sqrt(3.0) - 2.0
    assert np.allclose(h, np.sqrt(3.0) * z1 ** np.abs(np.arange(29) - 14), rtol=0, atol=2e-7)
    for x in range(lo, hi + 1):
        assert ops_.p_first[x] == x - c and np.array_equal(ops_.p_vals[x], h)
    small = rf.AxisOperators(24, 5, [0]).prefilter_toeplitz()           # too short for an interior
    assert small[2] > small[3] or small[3] - small[2] < 8


@pytest.mark.gpu
def test_prefilter_register_window_kernels_equal_table_kernels(cuda):
    """pcb_refi_prefilter_toeplitz == pcb_refi_prefilter bit for bit (same summation order)."""
    import ctypes as C
    import torch
    from plenopticam_b200 import _native as nat
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_refocuser import refinement as rf
    M, S, T = 3, 101, 118
    rng = np.random.default_rng(8)
    vp = ops.to_device(rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16))
    plan = rf.device_plan(S, T, M, [-1, 0, 1], np.zeros((M, M), dtype=bool), vp.device)
    (hx, cx, xlo, xhi), (hy, cy, ylo, yhi) = plan.tx, plan.ty
    assert hx is not None and hy is not None and xhi - xlo > 40 and yhi - ylo > 40
    L, p, st = nat.lib(), ops._ptr, ops._stream()
    outs = []
    for fast in (False, True):
        tmp = torch.zeros((M, M, S, T, 3), dtype=torch.float32, device=vp.device)
        coef = torch.zeros_like(tmp)
        if fast:
            nat.check(L.pcb_refi_prefilter_toeplitz(p(vp), ops._pcb_dtype(vp), M, S, T, 3, p(plan.py_v), p(plan.py_f),
                                                    plan.Kpy, p(plan.px_v), p(plan.px_f), plan.Kpx, p(plan.z_dev), p(tmp),
                                                    p(coef), hx.ctypes.data_as(C.c_void_p), int(hx.size), cx, xlo, xhi,
                                                    hy.ctypes.data_as(C.c_void_p), int(hy.size), cy, ylo, yhi, None, st))
        else:
            nat.check(L.pcb_refi_prefilter(p(vp), ops._pcb_dtype(vp), M, S, T, 3, p(plan.py_v), p(plan.py_f), plan.Kpy,
                                           p(plan.px_v), p(plan.px_f), plan.Kpx, p(plan.z_dev), p(tmp), p(coef), st))
        torch.cuda.synchronize()
        outs.append((tmp, coef))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])


def test_quad_tables_reproduce_the_bands():
    """Slice quads (refine_hq_kernel): the shared window + zero-padded float4 weights give the same sums as the plain
    per-slice bands, for a ragged number of slices, at the borders, and for masked / unmasked view columns."""
    from plenopticam_b200.lfp_refocuser import refinement as rf
    M, T = 9, 57
    a_list = list(range(-13, 10))                           # 23 slices: last quad has 3 live members
    c = M // 2
    Bx = rf.AxisOperators(T, M, sorted(set(a * d for a in a_list for d in range(-c, c + 1))))
    ix = np.array([[Bx.index[a * (c - i)] for i in range(M)] for a in a_list], dtype=np.int32)
    tab = rf.quad_tables(Bx, ix, M)
    assert tab is not None and tab["G"] == 6 and tab["W"].min() >= Bx.K and tab["W"].max() <= rf.QUAD_MAXW
    coef = np.random.default_rng(0).random(T)
    for n in (0, 5, 11, 22):
        g, q = divmod(n, rf.QUAD)
        for i in range(M):
            W = int(tab["W"][i])
            blk = tab["wq"][tab["woff"][i] * 4:(tab["woff"][i] + tab["G"] * W * T) * 4].reshape(tab["G"], W, T, 4)
            fu = tab["fu"][g, i]
            assert fu.min() >= 0 and (fu + W).max() <= T
            assert (fu - np.arange(T)).min() >= tab["lo"][i] and (fu - np.arange(T)).max() + W - 1 <= tab["hi"][i]
            out = np.array([sum(float(blk[g, t, x, q]) * coef[fu[x] + t] for t in range(W)) for x in range(T)])
            assert np.abs(out - Bx.apply_coef(a_list[n] * (c - i), coef)).max() < 1e-6
    # the padding members of the last quad carry zero weights
    W0 = int(tab["W"][0])
    last = tab["wq"][tab["woff"][0] * 4:(tab["woff"][0] + tab["G"] * W0 * T) * 4].reshape(tab["G"], W0, T, 4)[5]
    assert np.all(last[..., 3] == 0.0) and np.any(last[..., 2] != 0.0)


@pytest.mark.gpu
def test_quad_kernel_equals_tap_kernel_bit_for_bit(cuda):
    """refine_hq_kernel (slice quads) vs refine_h_kernel (one slice at a time): identical stacks."""
    import os
    import torch
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_refocuser.refinement import refocus_refi
    rng = np.random.default_rng(21)
    for M, S, T, a_list in ((7, 23, 140, list(range(-7, 8))), (15, 14, 300, list(range(-30, 30, 1))[:22])):
        vp = ops.to_device((rng.random((M, M, S, T, 3)) * 65535).astype(np.uint16))
        mask = ops.aperture_mask(M)
        got = refocus_refi(vp, a_list, mask)
        os.environ["PCB_REFINE_KERNEL"] = "taps"
        try:
            ref = refocus_refi(vp, a_list, mask)
        finally:
            os.environ.pop("PCB_REFINE_KERNEL", None)
        assert torch.equal(got, ref)


def test_upscaled_oracle_matches_the_reference_png_bytes():
    """The oracle's up-sampled sum -> auto_hist_align(ref = itself) -> sRGB -> uint8_norm equals, byte for byte, what the
    unmodified reference handed to imageio for its upscale_*.png files (oracle/gen_golden_upscale.py)."""
    g = golden("refocus_upscale_u16")
    ups = onp.refocus_refi_upscaled(g["vp"], tuple(int(v) for v in g["ran_refo"]), int(g["M"]))
    o8 = np.stack([onp.uint8_norm(onp.srgb_conv(onp.auto_hist_align(u, u))) for u in ups])
    assert np.array_equal(o8, g["upscaled_u8"])
    assert [str(n) for n in g["names"]] == ["%s.png" % np.round(a / 5, 2) for a in range(-5, 5)]


@pytest.mark.gpu
def test_upscale_sink_receives_the_reference_png_content(cuda):
    """LfpShiftAndSum.upscale_sink: the opt-in form of the reference's upscale_*.png side effect (lfp_shiftandsum.py:126-132).
    Quantised like misc.save_img_file would (Normalizer.uint8_norm) the images equal the reference's bytes up to 1 LSB, and
    the names are the reference's."""
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_refocuser import LfpShiftAndSum
    g = golden("refocus_upscale_u16")
    M = int(g["M"])
    cfg, sta = make_cfg(M=M, ran_refo=[int(v) for v in g["ran_refo"]], refi=True)
    got = []
    obj = LfpShiftAndSum(vp_img_arr=g["vp"], cfg=cfg, sta=sta)
    obj.upscale_sink = lambda fname, img: got.append((fname, np.array(img)))
    assert obj.main() is True
    assert np.abs(obj.refo_stack - g["stack"]).max() <= 5e-6 * np.abs(g["stack"]).max()
    assert ["%s.png" % f for f, _ in got] == [str(n) for n in g["names"]]
    for (_, img), want in zip(got, g["upscaled_u8"]):
        assert img.dtype == np.float32 and img.shape == want.shape
        q = ops.to_host(ops.uint8_norm(ops.to_device(img)))
        assert np.abs(q.astype(int) - want.astype(int)).max() <= 1


@pytest.mark.gpu
def test_upscaled_slices_vs_oracle(cuda):
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_refocuser.refinement import device_plan, upscaled_slices
    rng = np.random.default_rng(33)
    M, S, T = 7, 9, 12
    vp = (rng.random((M, M, S, T, 3)) * 65535).astype(np.uint16)
    ran = (-1, 1)
    a_all = list(range(M * ran[0], M * ran[1]))
    probe = [-7, -3, 0, 6]
    ref = onp.refocus_refi_upscaled(vp, ran, M, a_list=probe)
    dev = ops.to_device(vp)
    mask = ops.aperture_mask(M)
    plan = device_plan(S, T, M, a_all, mask, dev.device)
    for (a, U), want in zip(upscaled_slices(dev, probe, mask, plan, srgb=False), ref):
        got = ops.to_host(U)
        assert got.shape == (S * M, T * M, 3)
        assert np.abs(got - want).max() <= 5e-6 * np.abs(want).max(), a
