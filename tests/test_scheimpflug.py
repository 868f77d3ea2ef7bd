"""Scheimpflug composite ("next" row of SURVEY §8f): oracle vs golden vectors of the unmodified reference (CPU),
CUDA gather + uint16 normalisation vs the same vectors (GPU)."""
import os

import numpy as np
import pytest

from conftest import golden, make_cfg
from oracle import oracle_np as onp


def test_oracle_scheimpflug_equals_reference_golden():
    g = golden("scheimpflug_f32")
    for d in onp.PFLU_VALS:
        img, q = onp.scheimpflug_from_stack(g["stack"], d)
        assert np.array_equal(q, g[d.replace(' ', '_')]), d
    # the reference's `or` quirk: 'skew down' equals 'vertical', 'skew up' is the only diagonal map
    assert np.array_equal(g["skew_down"], g["vertical"]) and not np.array_equal(g["skew_up"], g["vertical"])


def test_scheimpflug_maps_and_mode():
    from plenopticam_b200 import ops
    a_y, a_x = ops.scheimpflug_maps(9, 23, 31)
    assert np.array_equal(a_x, np.linspace(0, 9, 33, dtype='int')[1:-1]) and a_y.size == 23
    assert [ops.scheimpflug_mode(d) for d in ops.PFLU_VALS] == [0, 1, 2, 0]


@pytest.mark.gpu
def test_scheimpflug_class_vs_reference_golden(cuda, tmp_path):
    from plenopticam_b200.lfp_refocuser import LfpScheimpflug
    g = golden("scheimpflug_f32")
    cfg, sta = make_cfg(M=7, ran_refo=(0, 9), tmpdir=tmp_path)
    obj = LfpScheimpflug(refo_stack=g["stack"], cfg=cfg, sta=sta)
    assert obj.main() is True
    for d in onp.PFLU_VALS:
        assert np.array_equal(obj.images_u16[d], g[d.replace(' ', '_')]), d       # bit-exact uint16
        ref_img, _ = onp.scheimpflug_from_stack(g["stack"], d)
        assert np.array_equal(obj.images[d], ref_img)                            # pure gather
        fn = os.path.join(obj.fp, 'scheimpflug_0_9_' + d + '.png')
        assert os.path.exists(fn)
    sta.interrupt = True
    assert LfpScheimpflug(refo_stack=g["stack"], cfg=cfg, sta=sta).main() is False


@pytest.mark.gpu
def test_refocuser_with_scheimpflug(cuda, tmp_path):
    from plenopticam_b200.lfp_refocuser import LfpRefocuser
    g = golden("refocuser_post_u16")
    cfg, sta = make_cfg(M=int(g["M"]), ran_refo=[int(v) for v in g["ran_refo"]], tmpdir=tmp_path)
    cfg.params[cfg.opt_refo] = True
    cfg.params[cfg.opt_pflu] = 'vertical'
    obj = LfpRefocuser(vp_img_arr=g["vp"], cfg=cfg, sta=sta)
    assert obj.main() is True
    ref, _ = onp.scheimpflug_from_stack(obj.refo_stack, 'horizontal')
    assert np.array_equal(obj.scheimpflug_images['horizontal'], ref)
