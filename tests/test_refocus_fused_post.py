"""pcb_refocus_int_srgb: the refocuser's colour management fused into the shift-and-sum epilogue must equal the
separate kernels bit for bit, and the reference within the north-star tolerance."""
import numpy as np
import pytest

from conftest import golden
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def _both(vp_np, a_list):
    import torch
    from plenopticam_b200 import ops
    vp = ops.to_device(vp_np)
    plan = ops.RefocusPlan(tuple(vp.shape), vp.dtype, a_list, None, vp.device)
    fused = ops.refocus_int_srgb(vp, a_list, plan=plan)
    assert plan.fused is True
    mm = ops.new_minmax()
    lin = ops.refocus_int(vp, a_list, mm=mm)
    sep = ops.refocuser_postprocess(lin, mm=mm)
    torch.cuda.synchronize()
    return fused, sep, lin


@pytest.mark.parametrize("M,S,T,a_list,seed", [
    (15, 37, 61, list(range(-32, 32)), 1),
    (7, 40, 50, [-2, -1, 0, 1], 2),
    (9, 21, 700, list(range(-3, 70)), 3),         # more slices than one control block holds (two launches per pass)
    (7, 5, 1100, [-3, 0, 2, 5], 4),               # two pixels per thread
])
def test_fused_equals_separate_and_oracle(cuda, M, S, T, a_list, seed):
    import torch
    rng = np.random.default_rng(seed)
    vp = rng.integers(0, 65536, size=(M, M, S, T, 3), dtype=np.uint16)
    fused, sep, lin = _both(vp, a_list)
    assert torch.equal(fused, sep)
    ref = onp.refocuser_postprocess(np.stack([onp.refocus_int_slice(onp._apply_aperture(vp.astype('float64'), M), a, M)
                                              for a in a_list[:3]]))
    lo, hi = onp.hist_percentiles(lin[0].cpu().numpy())
    exp = onp.srgb_conv(np.asarray(onp.norm_fun(lin[:3].cpu().numpy(), lo, hi), dtype='float32'))
    assert np.abs(fused[:3].cpu().numpy().astype(np.float64) - exp).max() <= 2e-6
    assert ref.shape == exp.shape


def test_fused_falsy_bound_takes_the_fix_up_pass(cuda):
    """Slice 0 almost black -> hi (99.9th percentile) == 0 -> Normalizer falls back to the stack maximum, which is
    known only after every slice is finalised: the fix-up launch has to rewrite the stack."""
    import torch
    M, S, T = 5, 40, 60
    vp = np.zeros((M, M, S, T, 3), dtype=np.uint16)
    vp[:, :, 3, 5, :] = 4000          # one bright pixel: 1 / 2400 of the samples of slice 0 < 0.1 % -> P99.9 == 0
    fused, sep, lin = _both(vp, [0, 1, 2])
    lo, hi = onp.hist_percentiles(lin[0].cpu().numpy())
    assert hi == 0.0 and lo == 0.0 and float(lin.max()) > 0
    assert torch.equal(fused, sep)
    exp = onp.refocuser_postprocess(lin.cpu().numpy())
    assert np.abs(fused.cpu().numpy().astype(np.float64) - exp).max() <= 2e-6
    assert float(fused.max()) > 0.5                                    # really normalised by the stack maximum


def test_fused_vs_reference_golden(cuda):
    from plenopticam_b200 import ops
    g = golden("refocuser_post_u16")
    a_list = list(range(*[int(v) for v in g["ran_refo"]]))
    out = ops.refocus_int_srgb(ops.to_device(g["vp"]), a_list)
    assert np.abs(out.cpu().numpy().astype(np.float64) - g["stack"]).max() <= 2e-6


def test_unsupported_light_field_falls_back_to_separate_kernels(cuda):
    """float32 light fields have no packed-pixel kernel: same API, separate kernels, same values."""
    import torch
    from plenopticam_b200 import ops
    rng = np.random.default_rng(5)
    vp = rng.random((5, 5, 12, 14, 3)).astype(np.float32)
    dev = ops.to_device(vp)
    plan = ops.RefocusPlan(tuple(dev.shape), dev.dtype, [-1, 0, 1], None, dev.device)
    out = ops.refocus_int_srgb(dev, [-1, 0, 1], plan=plan)
    assert plan.fused is False
    mm = ops.new_minmax()
    assert torch.equal(out, ops.refocuser_postprocess(ops.refocus_int(dev, [-1, 0, 1], mm=mm), mm=mm))


def test_pipeline_fused_equals_unfused(cuda):
    import torch
    from plenopticam_b200.pipeline import LightFieldPipeline
    cent = onp.synth_centroids(20, 24, 11.0, 'hex', jitter=0.1, origin=(9.0, 9.0), seed=3)
    H, W = int(cent[:, 0].max() + 10), int(cent[:, 1].max() + 10)
    raw = onp.synth_raw(H, W, 3, seed=4, dtype=np.uint16)
    outs = []
    for fuse in (True, False):
        pipe = LightFieldPipeline(cent, (H, W, 3), np.uint16, 9, 'hex', ran_refo=(-4, 4), fuse_post=fuse)
        outs.append(torch.from_numpy(pipe.run(raw).copy()))
        assert pipe.gather_fused is True                      # the stack is read straight from the aligned image
        assert pipe.refocus_plan_al.fused is (True if fuse else None)
    assert torch.equal(outs[0], outs[1])
