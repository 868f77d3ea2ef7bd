"""`LfpViewpoints` base (lfp_extractor/lfp_viewpoints.py:30-59,229-250): holds the 4-D light field
vp_img_arr[j, i, s, t, c].  The reference eagerly casts it to float64 on the host (:35); here the array is
kept in its own dtype (host ndarray or CUDA tensor) and uploaded once, because every consumer on the hot
path reads it on the device."""
import numpy as np

from .. import ops
from ..cfg import PlenopticamConfig
from ..misc import PlenopticamStatus


def is_device_tensor(x):
    return hasattr(x, 'is_cuda')


class LfpViewpoints(object):

    def __init__(self, *args, **kwargs):
        self._vp_img_arr = kwargs['vp_img_arr'] if 'vp_img_arr' in kwargs else None
        self.cfg = kwargs['cfg'] if 'cfg' in kwargs else PlenopticamConfig()
        self.sta = kwargs['sta'] if 'sta' in kwargs else PlenopticamStatus()
        self._size_pitch = self.cfg.params[self.cfg.ptc_leng]
        self._cent_pitch = self._size_pitch // 2
        if self._vp_img_arr is not None:
            shp = tuple(self._vp_img_arr.shape)
            self._DIMS = shp if len(shp) == 3 else shp + (1,)

    @property
    def vp_img_arr(self):
        return self._vp_img_arr

    @vp_img_arr.setter
    def vp_img_arr(self, vp_img_arr):
        self._vp_img_arr = vp_img_arr

    @property
    def central_view(self):
        if self._vp_img_arr is None:
            return None
        view = self._vp_img_arr[self._cent_pitch, self._cent_pitch, ...]
        return ops.to_host(view.contiguous()) if is_device_tensor(view) else np.array(view)

    def _vp_device(self):
        """The light field as a contiguous CUDA tensor (uploads a host array once)."""
        if self._vp_img_arr is None:
            return None
        if not is_device_tensor(self._vp_img_arr):
            arr = self._vp_img_arr
            if arr.ndim == 4:
                arr = arr[..., None]
            return ops.to_device(arr, exact_u16=True)
        return self._vp_img_arr

    def circular_view_aperture_mask(self):
        """Views the reference zeroes before refocusing (lfp_viewpoints.py:229-250, offset=0, no ellipse).
        The kernels skip these views instead of writing zeros into the light field."""
        return ops.aperture_mask(self._size_pitch)
