"""`CfaProcessor` white balance (lfp_aligner/cfa_processor.py:40-279): the element-wise half of the Bayer stage —
`safe_bayer_awb()` = highlight-protected white balance + soft clipping — on the device (csrc/cfa.cuh), with the
reference's constructor and the `set_gains` ordering.

The demosaic that `main()` runs next (`bay2rgb`, :73-94) is `colour_demosaicing.demosaicing_CFA_Bayer_Menon2007`,
a third-party package (requirements.txt: colour-demosaicing) that is neither part of the reference sources nor
installed here, so there is nothing to pin a restatement against: `main()` / `bay2rgb()` raise
NotImplementedError after the white balance has run (the balanced mosaic stays available as `bay_img`).
"""
import numpy as np
import torch

from .. import ops
from ..misc import PlenopticamStatus


class CfaProcessor(object):

    def __init__(self, bay_img=None, wht_img=None, cfg=None, sta=None):
        self.cfg = cfg
        self.sta = sta if sta is not None else PlenopticamStatus()
        self._bay_img = self._as_f32(bay_img)
        self._wht_img = self._as_f32(wht_img)
        lfpimg = getattr(cfg, 'lfpimg', None) or {}
        self._bit_pac = lfpimg['bit'] if 'bit' in lfpimg else 10
        self._gains = lfpimg['awb'] if 'awb' in lfpimg else [1, 1, 1, 1]
        self._bay_pat = lfpimg['bay'] if 'bay' in lfpimg else None
        self._lfpimg = lfpimg
        self._rgb_img = np.array([])
        self._dev_out = None

    @staticmethod
    def _as_f32(img):
        """astype('float32') of the reference constructor (:47-48); CUDA tensors stay on the device."""
        if img is None:
            return None
        if hasattr(img, 'is_cuda'):
            return img.to(torch.float32)
        return np.asarray(img).astype('float32') if isinstance(img, np.ndarray) else None

    def set_gains(self, gains=None):
        """Gains per Bayer-quad position (:257-279); None when the pattern / layout is not one the reference handles."""
        if gains is None or self._bay_img is None:
            return None
        shape = tuple(self._bay_img.shape)
        if len(shape) == 2 or (len(shape) == 3 and shape[-1] == 4):
            if self._bay_pat == "GRBG":
                self._gains = np.array([gains[2], gains[1], gains[0], gains[3]])
            elif self._bay_pat == "BGGR":
                self._gains = np.array([gains[0], gains[2], gains[3], gains[1]])
            else:
                return None
        elif len(shape) == 3 and shape[-1] == 3:
            self._gains = np.array([gains[1], (gains[2] + gains[3]) / 2, gains[0]])
        else:
            return None
        return self._gains

    def safe_bayer_awb(self):
        """White balance that keeps highlights (:233-250); the result (float64 mosaic) is `bay_img`."""
        if self._bay_img is None or len(self._bay_img.shape) != 2:
            # the reference indexes four mosaic channels (:192): only 2-D Bayer images get through it
            raise ValueError("CfaProcessor.safe_bayer_awb expects the 2-D Bayer mosaic")
        if self.set_gains(self._lfpimg['awb']) is None:
            raise ValueError("Bayer pattern %r is not handled by CfaProcessor.set_gains" % (self._bay_pat,))
        H, W = (int(v) for v in self._bay_img.shape)
        if H % 2 or W % 2:
            raise ValueError("Bayer mosaic needs even dimensions")
        bay = self._bay_img if hasattr(self._bay_img, 'is_cuda') else ops.to_device_any(self._bay_img)
        wht = self._wht_img
        if wht is not None and not hasattr(wht, 'is_cuda'):
            wht = ops.to_device_any(wht)
        exp_bias = self._lfpimg['exp'] if 'exp' in self._lfpimg else None
        self._dev_out = ops.cfa_safe_awb(bay.contiguous(), None if wht is None else wht.contiguous(), self._gains,
                                         exp_bias)
        return True

    def bay2rgb(self, method=2):
        raise NotImplementedError('demosaicing (colour_demosaicing, cfa_processor.py:73-94) is a third-party '
                                  'dependency outside the B200 hot path; use bay_img (balanced mosaic)')

    def main(self):
        if self.sta.interrupt:
            return False
        self.safe_bayer_awb()
        return self.bay2rgb(2)

    @property
    def bay_img(self):
        """Balanced mosaic, float64 (H, W) like the reference's `_bay_img` after safe_bayer_awb."""
        return ops.to_host(self._dev_out) if self._dev_out is not None else self._bay_img

    @property
    def bay_img_device(self):
        return self._dev_out

    @property
    def rgb_img(self):
        return self._rgb_img.copy()
