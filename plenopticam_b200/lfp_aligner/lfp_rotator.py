"""Drop-in `LfpRotator` (lfp_aligner/lfp_rotator.py:31-180): angle from an OLS fit of the middle lens row /
column (host, float64, a few hundred points), image rotation by the cubic-spline kernels, centroid rotation
on the host in float64."""
import numpy as np

from .. import ops
from ..cfg import PlenopticamConfig
from ..misc import PlenopticamStatus


def _ols_slope(t, y):
    """Slope of the least-squares line y ~ a + b t (same estimator as pinv([1 t]) @ y, lfp_rotator.py:103-118)."""
    t = np.asarray(t, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    tm, ym = t.mean(), y.mean()
    return float(np.dot(t - tm, y - ym) / np.dot(t - tm, t - tm))


def estimate_rad(centroids):
    """lfp_rotator.py:68-100: average of the row tilt and the (negated) column tilt; every other lens of the
    middle column is skipped to cancel the hexagonal zig-zag."""
    c = np.asarray(centroids, dtype=np.float64)
    mid_row = c[c[:, 2] == int(c[:, 2].max() / 2)][:, :2]
    mid_col = c[c[:, 3] == int(c[:, 3].max() / 2)][:, :2][0::2]
    slope_row = _ols_slope(mid_row[:, 1], mid_row[:, 0])
    slope_col = _ols_slope(mid_col[:, 0], mid_col[:, 1])
    return (np.arctan(slope_row) - np.arctan(slope_col)) / 2


def rotate_centroids(centroids, rad, shape):
    """lfp_rotator.py:147-172: counter-clockwise rotation of the (y, x) columns about (shape-1)/2."""
    c = np.array(centroids, dtype=np.float64)
    ctr = (np.asarray(shape[:2], dtype=np.float64) - 1) / 2.
    cs, sn = np.cos(rad), np.sin(rad)
    y, x = c[:, 0] - ctr[0], c[:, 1] - ctr[1]
    c[:, 0] = cs * y - sn * x + ctr[0]
    c[:, 1] = sn * y + cs * x + ctr[1]
    return c


class LfpRotator(object):

    def __init__(self, lfp_img, mic_list, rad=None, cfg=None, sta=None):
        self._lfp_img = lfp_img
        self.cfg = cfg if cfg is not None else PlenopticamConfig()
        self.sta = sta if sta is not None else PlenopticamStatus()
        self._centroids = np.asarray(mic_list, dtype=np.float64)
        self._rad = rad
        self._dev_img = None

    def main(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        if self._rad is None:
            self.sta.status_msg('Rotation angle estimation', prnt)
            self.sta.progress(None, prnt)
            self._rad = estimate_rad(self._centroids)
            self.sta.progress(100, prnt)

        self.sta.status_msg('Rotate image', prnt)
        self.sta.progress(None, prnt)
        if self._rad * 180 / np.pi != 0:
            img = self._lfp_img
            squeeze = img.ndim == 2
            dev = img if hasattr(img, 'is_cuda') else ops.to_device(img[..., None] if squeeze else img, exact_u16=True)
            self._dev_img = ops.rotate(dev, self._rad)
            self._squeeze = squeeze
        self.sta.progress(100, prnt)

        self.sta.status_msg('Rotate centroids', prnt)
        self.sta.progress(None, prnt)
        self._centroids = rotate_centroids(self._centroids, self._rad, self._lfp_img.shape)
        self.sta.progress(100, prnt)
        return True

    @property
    def lfp_img(self):
        """float64 host copy, like the reference's `ndimage.rotate` output."""
        if self._dev_img is not None:
            out = ops.to_host(self._dev_img).astype(np.float64)
            return out[..., 0] if self._squeeze else out
        return np.array(self._lfp_img)

    @property
    def lfp_img_device(self):
        """float32 CUDA tensor of the rotated image (None if the angle was exactly 0)."""
        return self._dev_img

    @property
    def centroids(self):
        return self._centroids.tolist()
