"""BASELINE.json's full sizes (Illum-shaped exposure: 5368 x 7728 x 3 raw, 434 x 541 hex lenslets, M = 15, 64 slices),
checked through properties that do not need the oracle to finish: convexity of the resampling, gather round trip and
checksum, known-answer slices of the integer stack, fused == separate epilogue."""
import argparse

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def illum(cuda):
    import torch
    import bench
    from plenopticam_b200.pipeline import LightFieldPipeline
    w = bench.workload(argparse.Namespace(small=False, slices=64))
    cent = bench.make_centroids(w)
    pipe = LightFieldPipeline(cent, (w["H"], w["W"], w["C"]), np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"])
    raw = bench.synth_raw_device(torch, w["H"], w["W"], w["C"], seed=77, device=pipe.device)
    return w, pipe, raw


def test_resampling_is_a_convex_combination_and_minmax_is_exact(illum):
    import torch
    from plenopticam_b200 import ops
    w, pipe, raw = illum
    al32, mm = ops.resample_f32(raw, pipe.tabs, want_minmax=True)
    assert tuple(al32.shape) == (w["LY"] * 15, 625 * 15, 3)
    r32 = raw.view(torch.int16).to(torch.int32) & 0xffff
    rmin, rmax = float(r32.min()), float(r32.max())
    del r32
    lo, hi = float(al32.min()), float(al32.max())
    assert lo >= rmin * (1 - 1e-6) - 1e-3 and hi <= rmax * (1 + 1e-6) + 1e-3      # bilinear + hex lerp: weights in [0, 1]
    assert ops.minmax_values(mm) == (lo, hi)                                       # fused min / max == the data's
    al16, _ = ops.resample_u16(raw, pipe.tabs)
    assert int(al16.view(torch.int16).to(torch.int32).bitwise_and(0xffff).max()) == 65535      # uint16_norm spans the range
    assert int(al16.view(torch.int16).to(torch.int32).bitwise_and(0xffff).min()) == 0


def test_skipping_minmax_pass_is_deterministic_at_full_size(illum):
    """Regression (round 2): the min/max pass drops tiles that cannot move the global bounds; the bounds it compares with
    change while the kernel runs, so every CTA must take ONE snapshot of them - with per-warp snapshots some warps of a CTA
    left and the block reduction folded shared memory they never wrote (garbage bounds, seen with these seeds)."""
    import torch
    import bench
    from plenopticam_b200 import ops
    w, pipe, _ = illum
    for seed in (77, 1):
        raw = bench.synth_raw_device(torch, w["H"], w["W"], w["C"], seed=seed, device=pipe.device)
        _, mm32 = ops.resample_f32(raw, pipe.tabs, want_minmax=True)
        want = ops.minmax_values(mm32)
        for _ in range(4):
            _, mm = ops.resample_u16(raw, pipe.tabs)
            assert ops.minmax_values(mm) == want


def test_gather_round_trip_and_checksum(illum):
    import torch
    from plenopticam_b200 import ops
    w, pipe, raw = illum
    aligned = pipe.align(raw)
    vp = ops.viewpoints(aligned, 15)
    assert tuple(vp.shape) == (15, 15, 434, 625, 3)
    assert torch.equal(ops.viewpoints_inverse(vp).view(torch.int16), aligned.view(torch.int16))
    s_al = (aligned.view(torch.int16).to(torch.int64) & 0xffff).sum()
    s_vp = (vp.view(torch.int16).to(torch.int64) & 0xffff).sum()
    assert int(s_al) == int(s_vp)
    # one view is the stride-15 sub-grid of the aligned image
    assert torch.equal(vp[3, 11].view(torch.int16), aligned[3::15, 11::15].contiguous().view(torch.int16))


def test_stack_known_answers(illum):
    import torch
    from plenopticam_b200 import ops
    w, pipe, raw = illum
    pipe.align(raw)
    vp = pipe.viewpoints()
    a_list = list(range(-32, 32))
    stack = ops.refocus_int(vp, a_list, plan=pipe.refocus_plan)
    mask = torch.from_numpy(~ops.aperture_mask(15)).to(vp.device)
    v64 = (vp.view(torch.int16).to(torch.int64) & 0xffff)
    # a = 0: no shift at all -> the plain sum of the 177 views inside the aperture, divided by M
    plain = (v64 * mask[:, :, None, None, None]).sum(dim=(0, 1))
    assert torch.equal(stack[a_list.index(0)], (plain.to(torch.float64) / 15).to(torch.float32))
    # the centre pixel region far from the border: slice a sums vp[j, i, y + a (7 - j), x + a (7 - i)]
    a, y, x = 5, 217, 312
    acc = 0
    for j in range(15):
        for i in range(15):
            if bool(mask[j, i]):
                acc += int(v64[j, i, y + a * (7 - j), x + a * (7 - i), 1])
    assert float(stack[a_list.index(a), y, x, 1]) == np.float32(acc / 15)
    # a constant light field gives a constant stack: c * 177 / 15 exactly
    vp.view(torch.int16).fill_(1000)
    const = ops.refocus_int(vp, a_list, plan=pipe.refocus_plan)
    assert float(const.min()) == float(const.max()) == float(np.float32(1000 * 177 / 15))


def test_fused_epilogue_equals_separate_kernels(illum):
    import torch
    w, pipe, raw = illum
    pipe.align(raw)
    pipe.viewpoints()
    fused = pipe.refocus_post(torch.empty_like(pipe.post)).clone()
    assert pipe.refocus_plan.fused is True
    pipe.refocus()
    sep = pipe.post_process(torch.empty_like(pipe.post))
    assert torch.equal(fused, sep)
    assert float(fused.min()) >= 0.0 and float(fused.max()) <= 1.0
