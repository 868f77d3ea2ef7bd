"""Drop-in `LfpResampler`: same constructor kwargs, `main()` return values, properties, status messages,
cfg side effects and `lfp_img_align.pkl` checkpoint as the reference
(lfp_aligner/lfp_resampler.py:11-78, lfp_aligner/lfp_local_resampler.py:11-148), with the per-lenslet
interp2d / np.interp loops replaced by two launches of the sm_100a resampling kernel.

`smp_meth='local'` is the centroid-based resampler of the hot path (`method='linear'`, the only method reachable
from LfpAligner, lfp_aligner/top_level.py:74); `smp_meth='global'` (the reference's default,
lfp_aligner/lfp_global_resampler.py) is the projective rectification + hexagonal row merge, planned on the host in
float64 (lfp_global_resampler.py here) and evaluated on the device.
"""
import os
import pickle

import numpy as np

from .. import ops
from ..cfg.cfg import SMPL_METH
from .lfp_microlenses import LfpMicroLenses, build_tables


class LfpResampler(LfpMicroLenses):

    #: write `<exp_path>/lfp_img_align.pkl` like the reference does (lfp_resampler.py:64-68)
    write_checkpoint = True

    def __init__(self, *args, **kwargs):
        super(LfpResampler, self).__init__(*args, **kwargs)
        method = kwargs['method'] if 'method' in kwargs else 'linear'
        if method not in ('linear', None):
            raise NotImplementedError("LfpResampler(method=%r): only 'linear' is implemented on the GPU path "
                                      "(lfp_local_resampler.py:16-27)" % (method,))
        self._method = 'linear'
        self._minmax = None          # (min, max) of the float32 aligned image, filled by main()
        self._dev_align = None       # uint16 CUDA tensor, kept for device-resident pipelines

    # -- the reference's public surface ----------------------------------------------------------------
    def main(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.status_msg('Light-field alignment', prnt)

        meth = self.cfg.params[self.cfg.smp_meth]
        if meth == 'global' or not meth:
            self.global_resampling()
        elif meth == 'local':
            self.local_resampling()
        if meth not in SMPL_METH and meth:
            self.sta.status_msg('Resampling method %s unrecognized.' % meth, prnt)
            self.sta.interrupt = True

        self._write_lfp_align()
        return True

    def local_resampling(self):
        """Launches the min/max pass and the recompute+quantise pass; result stays on the device."""
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        pat = self.cfg.calibs[self.cfg.pat_type]
        if pat not in ('rec', 'hex'):
            return True                       # the reference silently does nothing (lfp_local_resampler.py:37-40)
        img = self._lfp_img if self._lfp_img.ndim == 3 else None
        if img is None:
            # a 2-D image crashes the reference at lfp_local_resampler.py:54 (window[:, :, p])
            raise IndexError('too many indices for array: LfpResampler needs a (H, W, C) image')
        tables = build_tables(self._CENTROIDS, img.shape, self._size_pitch, pat, hex_odd=self._hex_odd)
        if tables['n_invalid']:
            self.sta.status_msg('Warning: chosen micro image size exceeds light-field borders', prnt)
        tabs = ops.LensTables(tables)
        raw = img if hasattr(img, 'is_cuda') else ops.to_device(img, exact_u16=True)
        self._dev_align, mm = ops.resample_u16(raw, tabs, flip=self._flip)
        self._mm_dev = mm
        self.sta.progress(100, prnt)
        return True

    def global_resampling(self):
        """LfpGlobalResampler.global_resampling (lfp_global_resampler.py:20-30): projective rectification of the
        whole exposure, then (hex) row merge / stretch / odd-pitch resize; quantised to uint16 on the device."""
        from . import lfp_global_resampler as glob
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        pat = self.cfg.calibs[self.cfg.pat_type]
        img = self._lfp_img
        if getattr(img, 'ndim', 0) != 3 and not (hasattr(img, 'dim') and img.dim() == 3):
            raise IndexError('LfpResampler needs a (H, W, C) image')
        tra, oshape, pitch = glob.projective_plan(self._CENTROIDS, tuple(img.shape), pat, self._hex_odd)
        self._limg_pitch = np.asarray(pitch).astype('int')
        raw = img if hasattr(img, 'is_cuda') else ops.to_device(img, exact_u16=True)
        prj = ops.warp_projective(raw, tra, oshape)
        if pat == 'hex':
            if 0 in self._limg_pitch:
                raise Exception('Micro image size cannot be zero.')
            plan = glob.hex_plan(pitch, self._LENS_Y_MAX, self._LENS_X_MAX, int(prj.shape[1]), self._hex_odd)
            prj = ops.hex_align(prj, plan, self._LENS_Y_MAX)
            self._limg_pitch = np.asarray(plan["lnew"]).astype('int')
        self._dev_align_f32 = prj
        mm = ops.minmax_f32(prj)
        self._dev_align = ops.quantize_u16(prj, mm)
        self._mm_dev = mm
        self.sta.progress(100, prnt)
        return True

    def _write_lfp_align(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.status_msg('Save aligned light-field', prnt)
        self.sta.progress(None, prnt)
        if self._dev_align is not None:
            self._lfp_img_align = ops.to_host(self._dev_align)      # uint16, already normalised on the device
            self._minmax = ops.minmax_values(self._mm_dev)
        if self.write_checkpoint and self._lfp_img_align is not None:
            os.makedirs(self.cfg.exp_path, exist_ok=True)
            with open(os.path.join(self.cfg.exp_path, 'lfp_img_align.pkl'), 'wb') as f:
                pickle.dump(np.asarray(self._lfp_img_align).view(np.ndarray), f)      # a plain ndarray, like the reference's
        self.sta.progress(100, prnt)
        return True

    # -- extras for device-resident pipelines ------------------------------------------------------------
    @property
    def lfp_img_align_device(self):
        """The uint16 aligned image as a CUDA tensor (no copy)."""
        return self._dev_align

    @property
    def minmax(self):
        return self._minmax
