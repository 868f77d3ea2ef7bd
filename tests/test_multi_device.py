"""A pipeline built for cuda:1 must run on cuda:1 whatever the caller's current device is: launches go to
the executor's own device and stream, grids are sized for that device.  Needs two GPUs (skipped on the one-GPU test box;
run with `gpurun --gpus 2`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pipeline_on_second_device_while_first_is_current(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from plenopticam_b200.misc import synth
    from plenopticam_b200.pipeline import LightFieldPipeline
    w = synth.illum_workload(7, small=True)
    cent = synth.illum_centroids(w)
    raw = synth.synth_raw(w["H"], w["W"], 3, seed=11)
    torch.cuda.set_device(0)
    ref = LightFieldPipeline(cent, raw.shape, np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"]).run(raw).copy()
    pipe1 = LightFieldPipeline(cent, raw.shape, np.uint16, w["M"], w["pat"], ran_refo=w["ran_refo"], device="cuda:1")
    assert torch.cuda.current_device() == 0
    out = pipe1.run(raw).copy()                               # current device is still cuda:0
    assert pipe1.aligned.device.index == 1 and pipe1.post.device.index == 1
    assert torch.cuda.current_device() == 0
    assert np.array_equal(out, ref)
    es = pipe1.exposure_stream(depth=2)
    es.pinned_input(0)[...] = raw
    k = es.submit()
    assert np.array_equal(es.result(k), ref)
    assert torch.cuda.current_device() == 0
