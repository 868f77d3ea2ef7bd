"""Drop-in `LfpCropper` (lfp_extractor/lfp_cropper.py:28-66): central crop of every micro image when the
requested angular size is smaller than the aligned one; pass-through when equal."""
import numpy as np

from .. import ops
from ..lfp_aligner.lfp_microlenses import LfpMicroLenses


class LfpCropper(LfpMicroLenses):

    def __init__(self, *args, **kwargs):
        super(LfpCropper, self).__init__(*args, **kwargs)
        self._k = (self._limg_pitch - self._size_pitch).astype('int') // 2
        self.new_lfp_img = None
        if self._lfp_img_align is not None:
            self._LENS_Y_MAX = int(self._lfp_img_align.shape[0] / self._limg_pitch[0])
            self._LENS_X_MAX = int(self._lfp_img_align.shape[1] / self._limg_pitch[1])

    def main(self):
        if self.sta.interrupt:
            return False
        M_out = self._size_pitch
        if M_out < self._limg_pitch[0] or M_out < self._limg_pitch[1]:
            if (self._limg_pitch[0] - M_out) % 2 or (self._limg_pitch[1] - M_out) % 2 or M_out > min(self._limg_pitch):
                # an odd margin makes the reference's slice assignment (:60-62) raise a broadcast ValueError
                raise ValueError('could not broadcast input array: micro images of size %s cannot be cropped to %d'
                                 % (tuple(int(v) for v in self._limg_pitch), M_out))
            self.sta.status_msg('Render angular domain', self.cfg.params[self.cfg.opt_prnt])
            src = self._lfp_img_align
            dev = src if hasattr(src, 'is_cuda') else ops.to_device(src if src.ndim == 3 else src[..., None])
            self._dev_out = ops.crop_micro_images(dev, (int(self._limg_pitch[0]), int(self._limg_pitch[1])), int(M_out))
            self.new_lfp_img = self._dev_out if hasattr(src, 'is_cuda') else ops.to_host(self._dev_out)
            self.sta.progress(100, self.cfg.params[self.cfg.opt_prnt])
        elif M_out == min(self._limg_pitch):
            self.new_lfp_img = self._lfp_img_align

    @property
    def lfp_img_align(self):
        if self.new_lfp_img is None and self._lfp_img_align is not None:
            # main() did not run or had nothing to do (requested size above the aligned one): the reference hands out
            # the zero image it pre-allocated in its constructor (lfp_cropper.py:37-43,65-66)
            src = self._lfp_img_align
            p = int(src.shape[-1]) if len(src.shape) == 3 else 1
            shp = (int(self._size_pitch * self._LENS_Y_MAX), int(self._size_pitch * self._LENS_X_MAX), p)
            if hasattr(src, 'is_cuda'):
                import torch
                return torch.zeros(shp, dtype=src.dtype, device=src.device)
            return np.zeros(shp, dtype=src.dtype)
        if self.new_lfp_img is None:
            return None
        return self.new_lfp_img          # handed over, not copied (the reference copies, lfp_cropper.py:65-66)
