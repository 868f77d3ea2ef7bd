"""Drop-in `LfpContrast` for the extractor's default colour management (lfp_extractor/lfp_contrast.py:31-67,
249-263): `main()` aligns the histogram of the whole 4-D light field to the central view
(`auto_hist_align(..., opt=True)`: lo = 0.005th percentile of its luma, hi = 99.9th percentile) and applies the
sRGB transfer curve — the same two kernels as the refocuser's post-process, reading the uint16 light field
directly (csrc/post.cuh).  Result: float32 in [0, 1] like the reference.

`opt_cont`, `opt_awb_` and `opt_sat_` (all off by default, cfg/constants.py:46-62) go through the third-party
`color_space_converter` (yuv / hsv conversions, source absent) and raise NotImplementedError.
"""
from .. import ops
from .lfp_viewpoints import LfpViewpoints, is_device_tensor


class LfpContrast(LfpViewpoints):

    def __init__(self, p_lo=None, p_hi=None, *args, **kwargs):
        super(LfpContrast, self).__init__(*args, **kwargs)
        self.p_lo = p_lo if p_lo is not None else 0.0
        self.p_hi = p_hi if p_hi is not None else 1.0
        self.ref_img = kwargs['ref_img'] if 'ref_img' in kwargs else None
        self._contrast, self._brightness = (1., 1.)
        self._dev_out = None

    def main(self):
        if self.sta.interrupt:
            return False
        for opt in ('opt_cont', 'opt_awb_', 'opt_sat_'):
            key = getattr(self.cfg, opt, opt)
            if self.cfg.params.get(key, False):
                raise NotImplementedError('LfpContrast %s (lfp_contrast.py:51-61) needs the third-party '
                                          'color_space_converter and is not part of the B200 path' % opt)
        vp = self._vp_device()
        c = self._cent_pitch
        self._dev_out = ops.viewpoints_contrast(vp, ref_view=vp[c, c])
        self._vp_img_arr = self._dev_out
        return True

    @property
    def vp_img_arr(self):
        arr = self._vp_img_arr
        return ops.to_host(arr) if is_device_tensor(arr) and arr is self._dev_out else arr

    @vp_img_arr.setter
    def vp_img_arr(self, vp_img_arr):
        self._vp_img_arr = vp_img_arr

    @property
    def vp_img_arr_device(self):
        return self._dev_out
