"""Drop-in `LfpDevignetter` (lfp_aligner/lfp_devignetter.py:33-183): same constructor kwargs, `main()` return
value, `lfp_img` / `wht_img` / `noise_lev` results, with the white-image conversion, the percentile
normalisation, the noise-level estimate and the division done on the device in float64 (bit-identical
divisions; csrc/devignette.cuh).

`patch_mode` is hard-wired to False in the reference (:43), so the per-lenslet polynomial fit
(`patch_devignetting`, `fit_patch`) is unreachable there and is not offered here.
"""
import numpy as np
import torch

from .. import ops
from .lfp_microlenses import LfpMicroLenses


def gauss_kernel_1d(length, sigma=1.0):
    """1-D factor of misc.create_gauss_kernel (misc/data_proc.py:31-42): the 2-D kernel is outer(g, g)."""
    length = (length - 1) // 2 + (length + 1) // 2
    xs = np.arange(-length // 2 + 1, length // 2 + 1)
    g = np.exp(-(xs.astype(np.float64) ** 2) / (2 * sigma ** 2))
    return g / g.sum()


class LfpDevignetter(LfpMicroLenses):

    def __init__(self, *args, **kwargs):
        super(LfpDevignetter, self).__init__(*args, **kwargs)
        self.noise_lev = kwargs['noise_lev'] if 'noise_lev' in kwargs else None
        self.noise_th = 0.05
        self.patch_mode = False
        self._dev_out = None         # float32 CUDA tensor for device-resident pipelines
        self._dev_wn = None
        self._out64 = None
        lfp = self._lfp_img
        if lfp is None:
            raise TypeError("LfpDevignetter needs lfp_img")
        wht = self._wht_img
        if wht is not None and hasattr(wht, 'shape') and not hasattr(wht, 'is_cuda'):
            wht = np.asarray(wht)
            if wht.ndim == 2 and lfp.ndim == 3:
                # the reference calls rgb2gry on the 2-D array here (:56), which mixes image columns
                raise ValueError("white image (H, W) with a (H, W, C) light-field image is not supported")
        self._wht_in = wht

    def main(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        lfp = self._lfp_img if hasattr(self._lfp_img, 'is_cuda') else ops.to_device_any(self._lfp_img)
        if self._wht_in is None:
            wht = torch.ones(tuple(lfp.shape), dtype=torch.float64, device=lfp.device)     # :36
        else:
            wht = self._wht_in if hasattr(self._wht_in, 'is_cuda') else ops.to_device_any(self._wht_in)
        self.sta.status_msg('Estimate white image noise level', prnt)
        user_pitch = np.mean(self.cfg.calibs[self.cfg.ptc_mean])
        out64, wn, stats = ops.devignette(lfp, wht, gauss_kernel_1d(int(user_pitch)), out_dtype=torch.float64)
        self.sta.status_msg('De-vignetting', prnt)
        self.sta.progress(None, prnt)
        self._out64, self._dev_wn = out64, wn
        self._wht_ndim3 = wht.dim() == 3
        if self.noise_lev is None:
            self.noise_lev = float(stats[2].item())
        self.sta.progress(100, prnt)
        return True

    # -- the reference's public surface ----------------------------------------------------------------
    @property
    def lfp_img(self):
        return ops.to_host(self._out64) if self._out64 is not None else super(LfpDevignetter, self).lfp_img

    @property
    def wht_img(self):
        if self._dev_wn is None:
            return self._wht_img
        w = ops.to_host(self._dev_wn)
        return w[..., None] if self._wht_ndim3 else w      # rgb2gry keeps a channel axis (H, W, 1)

    # -- extras for device-resident pipelines ------------------------------------------------------------
    @property
    def lfp_img_device(self):
        """De-vignetted image as a float32 CUDA tensor (what the resampling kernel reads)."""
        if self._dev_out is None and self._out64 is not None:
            self._dev_out = self._out64.to(torch.float32)
        return self._dev_out
