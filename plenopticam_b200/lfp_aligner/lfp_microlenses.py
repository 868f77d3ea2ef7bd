"""Host-side micro-lens bookkeeping: mirror of `LfpMicroLenses` (lfp_aligner/lfp_microlenses.py:30-253)
plus the float64 table preparation the resampling kernel consumes.

Everything here is tiny, per-camera (not per-pixel) work and stays on the host in float64 so that the
centroid rounding (`rint`, misc/type_checks.py:57-58) and `hex_odd` are bit-exact with the reference.
The reference looks a centroid up with a boolean mask over the whole table per lenslet (O(N^2),
lfp_microlenses.py:117); here the table is densified once.
"""
import os

import numpy as np

from .. import _native as nat
from ..cfg import PlenopticamConfig
from ..misc import PlenopticamStatus


# ------------------------------------------------------------------------------------------------------
# pure functions on the centroid table (rows [y, x, ly, lx])
# ------------------------------------------------------------------------------------------------------
def hex_direction(centroids):
    """True if the lens row under the upper-left lens is shifted to the right (lfp_microlenses.py:221-245)."""
    c = np.asarray(centroids, dtype=np.float64)
    first = c[(c[:, 2] == 0) & (c[:, 3] == 0)][:, :2].ravel()[:2]
    mid = int(c[:, 3].max() / 2)
    pitch = np.mean(np.diff(c[c[:, 3] == mid, 0]))
    below_left = ((c[:, 0] > first[0] + pitch / 2) & (c[:, 0] < first[0] + 3 * pitch / 2) &
                  (c[:, 1] < first[1]) & (c[:, 1] > first[1] - 3 * pitch / 4))
    return not bool(below_left.any())


def avg_pitch(centroids):
    """(pitch_y, pitch_x): ceil of the mean centroid spacing of the middle lens column / row
    (lfp_microlenses.py:159-174)."""
    c = np.asarray(centroids, dtype=np.float64)
    mid_row = int(c[:, 2].max() / 2)
    mid_col = int(c[:, 3].max() / 2)
    px = int(np.ceil(np.mean(np.diff(c[c[:, 2] == mid_row, 1]))))
    py = int(np.ceil(np.mean(np.diff(c[c[:, 3] == mid_col, 0]))))
    return np.array([py, px])


def dense_table(centroids):
    """(LY, LX, 2) float64 array of (y, x); first matching table row wins, a hole raises IndexError like
    `get_coords_by_idx` does (lfp_microlenses.py:113-119)."""
    c = np.asarray(centroids, dtype=np.float64)
    LY, LX = int(c[:, 2].max() + 1), int(c[:, 3].max() + 1)
    ly, lx = c[:, 2], c[:, 3]
    whole = (ly == np.floor(ly)) & (lx == np.floor(lx)) & (ly >= 0) & (lx >= 0)
    iy, ix = ly[whole].astype(np.int64), lx[whole].astype(np.int64)
    flat = iy * LX + ix
    # np.unique returns the index of the first occurrence of every lenslet
    uniq, first = np.unique(flat, return_index=True)
    if uniq.size != LY * LX:
        raise IndexError("centroid table has no entry for %d lenslet(s)" % (LY * LX - uniq.size))
    return c[whole][first, :2].reshape(LY, LX, 2)


def build_tables(centroids, shape, M, pat_type, hex_odd=None):
    """Prepare `pcb_lens_t[LY*LX]` and `pcb_hexcol_t[2*Xs]` (include/plenopticam_b200.h) in float64.

    Per lens (lfp_local_resampler.py:44-65,91-95): ry = rint(my), fy = my - ry; the bilinear sample for patch
    row v sits at ry - M//2 + v + fy, i.e. rows oy+v and oy+v+1 with oy = ry - M//2 + floor(fy) and weight
    wy = fy - floor(fy).  A (M+2)^2 window leaving the image, or an even M (window (M+3)^2), gives the
    reference's all-zero patch (:50,:58-60) -> PCB_LENS_INVALID.
    Per output lens column (hex only, :110,:129-136): xc = clip(linspace(0, LX, Xs)[k] + .5*parity, 0, LX-1).
    """
    H, W = int(shape[0]), int(shape[1])
    tab = dense_table(centroids)
    LY, LX = tab.shape[:2]
    c = M // 2
    r = np.rint(tab)                       # round-half-even on float64 == Python round()
    f = tab - r
    fl = np.floor(f)
    o = (r - c + fl).astype(np.int64)
    w = f - fl
    ri = r.astype(np.int64)
    valid = ((ri[..., 0] - c - 1 >= 0) & (ri[..., 0] + c + 2 <= H) &
             (ri[..., 1] - c - 1 >= 0) & (ri[..., 1] + c + 2 <= W))
    if M % 2 == 0:
        valid[...] = False
    lens = np.zeros(LY * LX, dtype=nat.LENS_DTYPE)
    lens["oy"] = np.where(valid, o[..., 0], nat.PCB_LENS_INVALID).ravel()
    lens["ox"] = np.where(valid, o[..., 1], 0).ravel()
    lens["wy"] = w[..., 0].ravel()
    lens["wx"] = w[..., 1].ravel()

    hexcol, Xs = None, LX
    if pat_type == 'hex':
        if hex_odd is None:
            hex_odd = hex_direction(centroids)
        Xs = int(np.round(2 * LX / np.sqrt(3)))
        hexcol = np.zeros(2 * Xs, dtype=nat.HEXCOL_DTYPE)
        base = np.linspace(0, LX, int(np.round(LX * 2 / np.sqrt(3))))
        for parity in (0, 1):
            xc = np.clip(base + .5 * parity, 0, LX - 1)
            x0 = np.floor(xc).astype(np.int64)
            sl = slice(parity * Xs, (parity + 1) * Xs)
            hexcol["x0"][sl] = x0
            hexcol["x1"][sl] = np.minimum(x0 + 1, LX - 1)
            hexcol["w"][sl] = xc - x0
    return dict(LY=LY, LX=LX, M=int(M), Xs=Xs, pat_type=pat_type, hex_odd=bool(hex_odd) if hex_odd is not None else False,
                lens=lens, hexcol=hexcol, n_invalid=int((~valid).sum()))


# ------------------------------------------------------------------------------------------------------
# class mirror
# ------------------------------------------------------------------------------------------------------
class LfpMicroLenses(object):
    """Same constructor kwargs, attributes and cfg side effects as the reference class
    (lfp_aligner/lfp_microlenses.py:32-80): `cfg`, `sta`, `lfp_img`, `wht_img`, `lfp_img_align`, `flip`.

    Difference: images are kept as given (no eager float64 host copies, :44-45); the `lfp_img` property
    still hands out a float64 copy."""

    def __init__(self, *args, **kwargs):
        self.cfg = kwargs['cfg'] if 'cfg' in kwargs else PlenopticamConfig()
        self.sta = kwargs['sta'] if 'sta' in kwargs else PlenopticamStatus()
        self._hex_odd = None
        self._lfp_img = kwargs.get('lfp_img')
        self._wht_img = kwargs.get('wht_img')
        self._lfp_img_align = kwargs.get('lfp_img_align')
        self._flip = kwargs.get('flip', False)

        if self.cfg.calibs:
            self._CENTROIDS = np.asarray(self.cfg.calibs[self.cfg.mic_list])
            self._LENS_Y_MAX = int(self._CENTROIDS[:, 2].max() + 1)
            self._LENS_X_MAX = int(self._CENTROIDS[:, 3].max() + 1)
            self._PAT_TYPE = self.cfg.calibs[self.cfg.pat_type] if self.cfg.pat_type in self.cfg.calibs else 'rec'
            self._hex_odd = hex_direction(self._CENTROIDS)

        self._size_pitch = 0
        self._cent_pitch = 0
        self._limg_pitch = self.lfp_align_pitch()
        self._cavg_pitch = self.centroid_avg_pitch()
        self._size_pitch = self.safe_pitch_eval(cavg_pitch=min(self._cavg_pitch),
                                                limg_pitch=min(self._limg_pitch),
                                                user_pitch=int(self.cfg.params[self.cfg.ptc_leng]))
        self._cent_pitch = self._size_pitch // 2

        if self._lfp_img is not None and hasattr(self._lfp_img, 'shape'):
            shp = tuple(self._lfp_img.shape)
            if len(shp) == 3:
                self._DIMS = shp
            elif len(shp) == 2:
                self._DIMS = shp + (1,)
            else:
                self.sta.status_msg('Incompatible image dimensions: Please either use KxLx3 or KxLx1 array dimensions')
                self.sta.error = True

    # -- pitch logic -----------------------------------------------------------------------------------
    def safe_pitch_eval(self, cavg_pitch, limg_pitch, user_pitch):
        """Decision table of lfp_microlenses.py:121-157 including its side effects on cfg / the pickle."""
        prnt = self.cfg.params[self.cfg.opt_prnt]
        safe = 3
        if 3 <= user_pitch <= cavg_pitch + 2:
            safe = user_pitch
        elif user_pitch > cavg_pitch:
            safe = cavg_pitch
            self.sta.status_msg('User size ({0} px) is larger than micro image size and reduced to {1} pixels.'
                                .format(user_pitch, cavg_pitch), prnt)
        elif user_pitch < 3 < cavg_pitch:
            safe = cavg_pitch
            self.sta.status_msg('User size ({0} px) is too small and increased to {1} pixels.'
                                .format(user_pitch, cavg_pitch), prnt)
        elif user_pitch < 3 and cavg_pitch < 3:
            self.sta.status_msg('Micro image dimensions are too small for light field computation.', True)
            self.sta.interrupt = True

        if 0 < limg_pitch < safe:
            # previous alignment was rendered with a smaller angular resolution: drop it and ask for a redo
            fp = os.path.join(self.cfg.exp_path, 'lfp_img_align.pkl')
            os.remove(fp)
            self.sta.status_msg('Angular resolution mismatch in previous alignment. Redo process')
            self.sta.interrupt = True
        elif limg_pitch >= safe > 0:
            self.cfg.params[self.cfg.ptc_leng] = safe
        elif limg_pitch == 0:
            self._size_pitch = safe
            self.cfg.params[self.cfg.ptc_leng] = safe
        return int(safe)

    def centroid_avg_pitch(self):
        if not hasattr(self, '_CENTROIDS'):
            return np.zeros(2)
        return avg_pitch(self._CENTROIDS)

    def lfp_align_pitch(self):
        """Micro-image size of an existing aligned image (lfp_microlenses.py:192-210)."""
        pitch = np.zeros(2, dtype='int')
        if self._lfp_img_align is None:
            return pitch
        rows, cols = self._lfp_img_align.shape[:2]
        if hasattr(self, '_LENS_Y_MAX') and hasattr(self, '_LENS_X_MAX'):
            h = np.sqrt(3) / 2 if self._PAT_TYPE == 'hex' else 1
            pitch[0] = int(round(rows / self._LENS_Y_MAX))
            pitch[1] = int(round(cols / self._LENS_X_MAX * h))
        else:
            pitch[0] = self.remainder_ratio(rows)
            pitch[1] = self.remainder_ratio(cols)
        return pitch.astype('int')

    @staticmethod
    def remainder_ratio(num):
        for d in range(3, 151):
            if num % d == 0:
                return int(d)
        return 0

    # -- lookups -----------------------------------------------------------------------------------------
    def get_coords_by_idx(self, ly, lx):
        hit = self._CENTROIDS[(self._CENTROIDS[:, 2] == ly) & (self._CENTROIDS[:, 3] == lx)]
        return hit[0, 0], hit[0, 1]

    @staticmethod
    def get_hex_direction(centroids):
        return hex_direction(centroids)

    def proc_lens_iter(self, fun, **kwargs):
        """Per-lenslet host callback loop (lfp_microlenses.py:82-111); kept for API completeness."""
        if kwargs.get('prnt', True):
            self.sta.status_msg(kwargs.get('msg', 'Light-field alignment process'), self.cfg.params[self.cfg.opt_prnt])
        extra = [kwargs[k] for k in kwargs if k not in ('cfg', 'sta', 'msg', 'prnt')]
        for ly in range(self._LENS_Y_MAX):
            for lx in range(self._LENS_X_MAX):
                fun(ly, lx, *extra)
            self.sta.progress((ly + 1) / self._LENS_Y_MAX * 100, self.cfg.params[self.cfg.opt_prnt])
            if self.sta.interrupt:
                return False
        return True

    @property
    def lfp_img(self):
        return np.array(self._lfp_img, dtype='float64') if self._lfp_img is not None else False

    @property
    def lfp_img_align(self):
        return self._lfp_img_align.copy() if self._lfp_img_align is not None else None
