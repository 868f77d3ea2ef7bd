"""Drop-in `LfpScheimpflug` for the stack-based composite (lfp_refocuser/lfp_scheimpflug.py:33-114): same
constructor, `main()` loops over the four directions and writes one PNG per direction into
`<lfp_path without extension>/scheimpflug/`.  The per-pixel Python loop is one gather kernel, the uint16
normalisation runs on the device as well.

Differences: the PNG is written with PIL as the 8-bit image `Normalizer(uint16 img).uint8_norm()` that
`misc.save_img_file` produces for .png (misc/file_rw.py:83-86) but WITHOUT the reference's 16x16 watermark
(`place_dnp`, file export is outside SURVEY §8); `scheimpflug_from_scratch` (a `pragma: no cover` alternate working
on the aligned image) is not offered.
"""
import os

import numpy as np

from .. import ops


class LfpScheimpflug(object):

    def __init__(self, refo_stack=None, lfp_img=None, cfg=None, sta=None):
        self.refo_stack = refo_stack
        self.lfp_img = lfp_img
        self.cfg = cfg
        self.sta = sta
        self.fp = os.path.join(os.path.splitext(self.cfg.params[self.cfg.lfp_path])[0], 'scheimpflug')
        self.images = {}              # direction -> float composite (numpy), like the reference's scheimpflug_img
        self.images_u16 = {}          # direction -> Normalizer(...).uint16_norm()
        self.write_files = True

    def main(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.status_msg('Scheimpflug extraction \n', prnt)
        self.sta.progress(None, prnt)
        os.makedirs(self.fp, exist_ok=True)
        for direction in list(ops.PFLU_VALS):
            self.cfg.params[self.cfg.opt_pflu] = direction
            self.sta.status_msg(direction + ':')
            if self.refo_stack is not None:
                if self.scheimpflug_from_stack() is False:
                    return False
            elif self.lfp_img is not None:
                raise NotImplementedError('scheimpflug_from_scratch (lfp_scheimpflug.py:116-200, pragma: no cover) '
                                          'is not part of the B200 path; pass refo_stack')
        return True

    def scheimpflug_from_stack(self):
        import torch
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        direction = self.cfg.params[self.cfg.opt_pflu]
        stack = self.refo_stack
        if not hasattr(stack, 'is_cuda'):
            arr = np.asarray(stack, dtype=np.float32)
            if arr.ndim == 3:
                arr = arr[..., None]
            stack = ops.to_device(np.ascontiguousarray(arr), dtype=np.float32)
        elif stack.dtype != torch.float32:
            stack = stack.to(torch.float32)
        img, img16 = ops.scheimpflug(stack, direction)
        self.images[direction] = ops.to_host(img).astype(np.float64)
        self.images_u16[direction] = ops.to_host(img16)
        self.sta.progress(100, prnt)
        if self.write_files:
            a_ran = self.cfg.params[self.cfg.ran_refo]
            fn = 'scheimpflug_' + str(a_ran[0]) + '_' + str(a_ran[-1]) + '_' + direction + '.png'
            self._save_png(self.images_u16[direction], os.path.join(self.fp, fn))
        return True

    @staticmethod
    def _save_png(img16, path):
        """8-bit PNG of an already uint16-normalised image: Normalizer(img).uint8_norm() (misc/normalizer.py:48-51)."""
        from PIL import Image
        x = img16.astype(np.float32)
        mn, mx = x.min(), x.max()
        x = (x - mn) / (mx - mn) if mx != mn and mx != 0 else x
        x8 = np.asarray(np.round(np.clip(x, 0, 1) * 255), dtype=np.uint8)
        Image.fromarray(x8[..., 0] if x8.shape[-1] == 1 else x8).save(path, 'PNG')
