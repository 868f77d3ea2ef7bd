"""Round-2 kernel forms, each checked against the form it replaces (which is itself pinned to the oracle / goldens):
the refocus kernel reading the aligned image directly (viewpoint gather + central crop folded into its staging) and
the min/max pass of the resampler that skips tiles which cannot move the global bounds."""
import os

import numpy as np
import pytest

from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def _aligned(cuda, Ma, S, T, seed):
    from plenopticam_b200 import ops
    rng = np.random.default_rng(seed)
    return ops.to_device(rng.integers(0, 65536, size=(S * Ma, T * Ma, 3), dtype=np.uint16))


@pytest.mark.parametrize("Ma,M,S,T,ran", [(7, 7, 12, 14, (-3, 4)), (9, 7, 17, 23, (-5, 3)), (15, 15, 20, 31, (-32, 32)),
                                          (15, 11, 9, 40, (-2, 9)), (5, 5, 33, 300, (-70, 71))])
def test_refocus_from_aligned_equals_gather_then_refocus(cuda, Ma, M, S, T, ran):
    import torch
    from plenopticam_b200 import ops
    al = _aligned(cuda, Ma, S, T, Ma * 100 + M)
    a_list = list(range(*ran))
    vp = ops.viewpoints(al, Ma, M)                       # LfpCropper + LfpRearranger
    ref = ops.refocus_int(vp, a_list)
    mm_ref = ops.new_minmax()
    ops.refocus_int(vp, a_list, mm=mm_ref)
    mm = ops.new_minmax()
    got = ops.refocus_int_from_aligned(al, Ma, a_list, M=M, mm=mm)
    assert torch.equal(got, ref)
    assert ops.minmax_values(mm) == ops.minmax_values(mm_ref)
    # and against the oracle on one slice
    vp64 = onp._apply_aperture(ops.to_host(vp), M)
    vp64 /= M
    a = a_list[len(a_list) // 3]
    assert np.array_equal(ops.to_host(got[a_list.index(a)]), onp.refocus_int_slice(vp64, a, M).astype(np.float32))


def test_refocus_from_aligned_full_mask_and_unaligned_base(cuda):
    """Every view active and the aligned image sub-allocated at odd 2-byte offsets: the first / last aligned rows share
    their 16-byte lines with memory outside the buffer (clamped bulk copies + plain edge loads)."""
    import torch
    from plenopticam_b200 import ops
    Ma, S, T = 5, 7, 13
    n = S * Ma * T * Ma * 3
    rng = np.random.default_rng(3)
    host = rng.integers(0, 65536, size=(S * Ma, T * Ma, 3), dtype=np.uint16)
    pool = torch.full((n + 64,), -1, dtype=torch.int16, device=cuda)
    full = np.zeros((Ma, Ma), dtype=bool)
    ref = None
    for off in (0, 1, 5, 8, 11):
        al = pool[off:off + n].view(torch.uint16).view(S * Ma, T * Ma, 3)
        al.view(torch.int16).copy_(torch.from_numpy(host.view(np.int16)).to(cuda))
        got = ops.refocus_int_from_aligned(al, Ma, range(-4, 5), view_zeroed=full)
        if ref is None:
            ref = ops.refocus_int(ops.viewpoints(al.clone(), Ma), range(-4, 5), view_zeroed=full)
        assert torch.equal(got, ref), off


def test_pipeline_gather_fused_equals_explicit_gather(cuda):
    import torch
    from plenopticam_b200.misc import synth
    from plenopticam_b200.pipeline import LightFieldPipeline
    w = synth.illum_workload(9, small=True)
    cent = synth.illum_centroids(w)
    raw = synth.synth_raw(w["H"], w["W"], 3, seed=4)
    outs = []
    for fused_gather in (True, False):
        for fuse_post in (True, False):
            pipe = LightFieldPipeline(cent, raw.shape, np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"],
                                      gather_fused=fused_gather, fuse_post=fuse_post)
            outs.append(torch.from_numpy(pipe.run(raw).copy()))
            assert pipe.gather_fused is fused_gather
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("pat,pad_w", [("hex", True), ("rec", True), ("hex", False)])
def test_minmax_pass_with_tile_skipping_is_exact(cuda, pat, pad_w):
    """The forms of the min/max pass - every tile interpolated; tiles that cannot move the bounds dropped after a scan of
    their raw samples, the box fetched by one tensor copy (needs a 16-byte row pitch) or by one bulk copy per row - give
    the bounds of the float32 image, bit for bit; so does the quantised image of the second pass."""
    import torch
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    from plenopticam_b200.misc import synth
    cent = synth.synth_centroids(40, 90, 14.25, pat, jitter=0.15, origin=(9.5, 9.5), seed=8)
    H, W = int(cent[:, 0].max() + 11), int(cent[:, 1].max() + 11)
    W = (W + 7) // 8 * 8 if pad_w else W // 8 * 8 + 3        # row pitch a multiple of 16 bytes (tensor copies) or not
    tabs = ops.LensTables(build_tables(cent, (H, W, 3), 15, pat))
    raws = [synth.synth_raw(H, W, 3, seed=9), np.full((H, W, 3), 65535, dtype=np.uint16)]
    raws.append(raws[0].copy())
    raws[2][7, 11] = 0                               # a single black sample inside the very first tile
    raws[2][H // 2, W // 2] = 65535
    raws.append(np.zeros((H, W, 3), dtype=np.uint16))
    raws.append((raws[0] >> 9).astype(np.uint16))    # few distinct values: many ties with the running bounds
    for raw in raws:
        dev = ops.to_device(raw)
        al32, mm32 = ops.resample_f32(dev, tabs, want_minmax=True)
        res, imgs = [], []
        for env in ({"PCB_RESAMPLE_NOSKIP": "1"}, {"PCB_RESAMPLE_ROWCOPIES": "1"}, {}):
            os.environ.update(env)
            try:
                al16, mm = ops.resample_u16(dev, tabs)
                res.append(ops.minmax_values(mm))
                imgs.append(al16)
            finally:
                for k in env:
                    os.environ.pop(k, None)
        assert all(r == res[0] for r in res), res
        assert res[0] == ops.minmax_values(mm32) == (float(al32.min()), float(al32.max()))
        assert all(torch.equal(i, imgs[0]) for i in imgs)
        assert torch.equal(imgs[0], ops.quantize_u16(al32, mm32))


def test_resampler_many_launches_on_two_streams(cuda):
    """Every launch carries its own tensor map (a kernel parameter): many launches in flight on two streams."""
    import torch
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    from plenopticam_b200.misc import synth
    cent = synth.synth_centroids(24, 50, 14.25, "hex", jitter=0.15, origin=(9.5, 9.5), seed=3)
    H, W = int(cent[:, 0].max() + 11), (int(cent[:, 1].max() + 11) + 7) // 8 * 8
    tabs = ops.LensTables(build_tables(cent, (H, W, 3), 15, "hex"))
    devs = [ops.to_device(synth.synth_raw(H, W, 3, seed=k)) for k in (2, 3)]
    refs = [ops.resample_u16(d, tabs)[0] for d in devs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda), torch.cuda.Stream(device=cuda)]
    outs = []
    for k in range(60):
        with torch.cuda.stream(streams[k & 1]):
            outs.append(ops.resample_u16(devs[k & 1], tabs)[0])
    torch.cuda.synchronize()
    assert all(torch.equal(o, refs[k & 1]) for k, o in enumerate(outs))


def _srgb_from_aligned(plan, al, pscratch):
    import torch
    from plenopticam_b200 import ops
    M, S, T, Cn = plan.shape
    out = torch.empty((plan.N, S, T, Cn), dtype=torch.float32, device=al.device)
    mm = ops.new_minmax()
    lohi = torch.empty(2, dtype=torch.float64, device=al.device)
    assert plan.launch_srgb(al, out, mm, lohi, pscratch) is True
    return out, ops.minmax_values(mm), lohi.cpu().numpy().copy()


def test_accumulators_come_back_zeroed_across_a_sequence_of_exposures(cuda, monkeypatch):
    """A plan that serves a SEQUENCE of exposures skips the accumulator memset from the second call on (the finalize pass
    stores zero behind every word it reads, PCB_REFOCUS_LEAVE_ZEROED / _SCRATCH_ZEROED): every stack must equal the one a
    fresh plan gives, for the linear and the fused sRGB form, incl. the degenerate exposures whose upper contrast bound
    is exactly 0 (then the fix-up pass re-reads the sums: the zeroing moves there) and an all-zero exposure."""
    import torch
    from plenopticam_b200 import _native as nat, ops
    monkeypatch.delenv("PCB_REFOCUS_MEMSET", raising=False)          # the A/B knob that forces the memset form
    Ma, M, S, T = 9, 7, 40, 52
    a_list = list(range(-4, 5))
    rng = np.random.default_rng(11)
    dense = [rng.integers(0, 65536, size=(S * Ma, T * Ma, 3), dtype=np.uint16) for _ in range(3)]
    spike = np.zeros_like(dense[0])                       # one bright sample: 99.9 % of slice 0 is zero, the stack is not
    spike[(S // 2) * Ma + Ma // 2, (T // 2) * Ma + Ma // 2, 1] = 60000
    dark = (dense[1] >> 12).astype(np.uint16)             # lower bound 0 (many zero samples), upper bound > 0
    zero = np.zeros_like(dense[0])
    seq = [dense[0], spike, dense[1], zero, dark, spike, spike, dense[2], zero, zero, dense[0]]
    pscratch = torch.empty(int(nat.lib().pcb_percentile_scratch_bytes()), dtype=torch.uint8, device=cuda)

    def plan():
        return ops.RefocusPlan((M, M, S, T, 3), torch.uint16, a_list, None, cuda, aligned_pitch=Ma)

    shared = plan()
    flags_seen = []
    for k, host in enumerate(seq):
        al = ops.to_device(host)
        was_clean = bool(getattr(shared.scratch, "_pcb_acc_zero", False))
        flags_seen.append(was_clean)
        if k % 3 == 2:                                     # linear form
            got = ops.refocus_int_from_aligned(al, Ma, a_list, M=M, plan=shared)
            ref = ops.refocus_int_from_aligned(al, Ma, a_list, M=M, plan=plan())
            assert torch.equal(got, ref), k
        else:                                              # fused colour management
            got, mm, lohi = _srgb_from_aligned(shared, al, pscratch)
            ref, mm_ref, lohi_ref = _srgb_from_aligned(plan(), al, pscratch)
            assert torch.equal(got, ref), k
            assert mm == mm_ref and np.array_equal(lohi, lohi_ref), k
            # ... and the separate kernels on the explicit gather (no flags anywhere on that route)
            lin = ops.refocus_int(ops.viewpoints(al, Ma, M), a_list)
            assert torch.equal(got, ops.refocuser_postprocess(lin)), k
        torch.cuda.synchronize()
        acc_bytes = len(a_list) * S * ((T + 3) // 4 * 4) * 12
        assert int(shared.scratch[:acc_bytes].count_nonzero()) == 0, k     # what the next call relies on
    assert flags_seen[0] is False and all(flags_seen[1:])


def test_a_launch_without_flags_marks_the_shared_scratch_dirty(cuda):
    """Plans of one geometry share the accumulators (LightFieldPipeline): the 4-D form leaves them dirty, so the aligned
    form that follows must zero them itself."""
    import torch
    from plenopticam_b200 import ops
    Ma, M, S, T = 7, 7, 12, 20
    a_list = list(range(-3, 4))
    al = _aligned(cuda, Ma, S, T, 5)
    p_al = ops.RefocusPlan((M, M, S, T, 3), torch.uint16, a_list, None, cuda, aligned_pitch=Ma)
    p_vp = ops.RefocusPlan((M, M, S, T, 3), torch.uint16, a_list, None, cuda)
    p_vp.scratch = p_al.scratch
    ref = ops.refocus_int_from_aligned(al, Ma, a_list, plan=p_al)
    assert p_al.scratch._pcb_acc_zero is True
    vp = ops.viewpoints(al, Ma)
    assert torch.equal(ops.refocus_int(vp, a_list, plan=p_vp), ref)
    assert p_al.scratch._pcb_acc_zero is False
    assert torch.equal(ops.refocus_int_from_aligned(al, Ma, a_list, plan=p_al), ref)
    assert torch.equal(ops.refocus_int_from_aligned(al, Ma, a_list, plan=p_al), ref)


@pytest.mark.parametrize("pat,W4,negative", [("hex", True, False), ("rec", True, False), ("hex", False, False),
                                             ("hex", True, True)])
def test_float32_exposures_through_the_pipelined_resampler(cuda, pat, W4, negative):
    """float32 exposures (what LfpAligner hands on after de-vignetting or rotation) take the persistent kernel too when
    the row pitch is a multiple of 16 bytes (W % 4 == 0: the ring slot is the tile, read in place), else the general
    tiled kernel.  Same image as the uint16 exposure of the same values, the bounds of the min/max pass (with tile
    skipping, also with negative samples) are those of the image, the quantised image is the quantised float image."""
    import torch
    from plenopticam_b200 import ops
    from plenopticam_b200.lfp_aligner import build_tables
    from plenopticam_b200.misc import synth
    cent = synth.synth_centroids(40, 90, 14.25, pat, jitter=0.15, origin=(9.5, 9.5), seed=8)
    H, W = int(cent[:, 0].max() + 11), int(cent[:, 1].max() + 11)
    W = (W + 3) // 4 * 4 if W4 else W // 4 * 4 + 1
    tabs = ops.LensTables(build_tables(cent, (H, W, 3), 15, pat))
    raw16 = synth.synth_raw(H, W, 3, seed=9)
    raw32 = raw16.astype(np.float32)
    if negative:
        raw32 = raw32 - 30000.0
    d16, d32 = ops.to_device(raw16), ops.to_device(raw32)
    al_ref = ops.resample_f32(d16, tabs)                              # uint16 exposure, persistent kernel (pinned elsewhere)
    al32, mm32 = ops.resample_f32(d32, tabs, want_minmax=True)
    off = 30000.0 if negative else 0.0
    assert float((al32 + off - al_ref).abs().max()) <= (2e-6 * 65535 if not negative else 2e-2)
    assert ops.minmax_values(mm32) == (float(al32.min()), float(al32.max()))
    q, mm = ops.resample_u16(d32, tabs)
    assert ops.minmax_values(mm) == ops.minmax_values(mm32)
    assert torch.equal(q, ops.quantize_u16(al32, mm32))
    os.environ["PCB_RESAMPLE_KERNEL"] = "tile"                       # the general kernel on the same exposure
    try:
        al_t, mm_t = ops.resample_f32(d32, tabs, want_minmax=True)
    finally:
        os.environ.pop("PCB_RESAMPLE_KERNEL", None)
    assert float((al32 - al_t).abs().max()) <= 2e-6 * 65535


def test_captured_exposure_replays_like_the_stream_launches(cuda):
    """LightFieldPipeline.capture: the launches of one exposure as a CUDA graph; replays on new contents of the bound raw
    tensor equal run_device, also after a launch that left the shared refocus scratch dirty."""
    import torch
    from plenopticam_b200 import ops
    from plenopticam_b200.misc import synth
    from plenopticam_b200.pipeline import LightFieldPipeline
    w = synth.illum_workload(9, small=True)
    cent = synth.illum_centroids(w)
    raws = [ops.to_device(synth.synth_raw(w["H"], w["W"], 3, seed=s)) for s in (4, 5, 6)]
    pipe = LightFieldPipeline(cent, tuple(raws[0].shape), np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"])
    refs = [pipe.run_device(r).clone() for r in raws]
    buf = raws[0].clone()
    step = pipe.capture(buf)
    for k in (1, 2, 0, 2):
        buf.copy_(raws[k])
        assert torch.equal(step.replay(), refs[k]), k
    pipe.viewpoints()
    pipe.refocus()                                       # the 4-D form: leaves the shared accumulators dirty
    assert pipe.refocus_plan.scratch._pcb_acc_zero is False
    buf.copy_(raws[1])
    assert torch.equal(step.replay(), refs[1])
    assert torch.equal(pipe.run_device(raws[2]), refs[2])              # and stream launches after a replay
