"""Drop-in `LfpRearranger` (lfp_extractor/lfp_rearranger.py:29-137): aligned micro-image array <-> 4-D
viewpoint array, one kernel launch each way instead of M*M strided numpy copies.  Bit-exact, dtype kept.
Host ndarrays in -> host ndarrays out; CUDA tensors in -> CUDA tensors out (no PCIe traffic)."""
import numpy as np

from .. import ops
from ..misc import PlenopticamError
from .lfp_viewpoints import LfpViewpoints, is_device_tensor


class LfpRearranger(LfpViewpoints):

    def __init__(self, lfp_img_align=None, *args, **kwargs):
        super(LfpRearranger, self).__init__(*args, **kwargs)
        self._lfp_img_align = lfp_img_align if lfp_img_align is not None else None
        src = self._lfp_img_align if self._lfp_img_align is not None else self._vp_img_arr
        self._dtype = src.dtype if src is not None else None

    def main(self):
        if self.sta.interrupt:
            return False
        self.compose_viewpoints()          # returns None like the reference (:70-77)

    def compose_viewpoints(self):
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.status_msg('Viewpoint composition', prnt)
        self.sta.progress(None, prnt)
        src = self._lfp_img_align
        if len(src.shape) not in (2, 3):
            raise PlenopticamError('Dimensions %s of provided light-field not supported', src.shape,
                                   cfg=self.cfg, sta=self.sta)
        if self.sta.interrupt:
            return False
        on_dev = is_device_tensor(src)
        if len(src.shape) == 2:
            src = src[..., None]
        dev = src if on_dev else ops.to_device(src)
        vp = ops.viewpoints(dev.contiguous() if on_dev else dev, int(self._size_pitch))
        self._dev_vp = vp                                  # kept for the stages that follow on the device
        self._vp_img_arr = vp if on_dev else ops.to_host(vp)
        self.sta.progress(100, prnt)
        return True

    def decompose_viewpoints(self):
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.status_msg('Viewpoint decomposition', prnt)
        self.sta.progress(None, prnt)
        vp = self._vp_img_arr
        if len(vp.shape) not in (4, 5):
            raise PlenopticamError('Dimensions %s of provided light-field not supported', vp.shape,
                                   cfg=self.cfg, sta=self.sta)
        if self.sta.interrupt:
            return False
        on_dev = is_device_tensor(vp)
        if len(vp.shape) == 4:
            vp = vp[..., None]
        if vp.shape[0] != vp.shape[1]:
            raise ValueError('square angular resolution expected, got %dx%d' % (vp.shape[0], vp.shape[1]))
        dev = vp if on_dev else ops.to_device(vp)
        out = ops.viewpoints_inverse(dev.contiguous() if on_dev else dev)
        self._lfp_img_align = out if on_dev else ops.to_host(out)
        self._size_pitch = vp.shape[0]
        self.sta.progress(100, prnt)
        return True

    @property
    def lfp_img_align(self):
        return self._lfp_img_align
