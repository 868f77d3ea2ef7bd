"""Drop-in `LfpShiftAndSum` (lfp_refocuser/lfp_shiftandsum.py:34-139,306-322): the whole refocus stack in one
kernel launch (all slices, all views, aperture mask and the /M scaling fused) instead of
N * M*M `np.pad` + `np.add` passes.

`refo_stack` is float64 (N,S,T,C) like the reference's; the device result is float32 and is exposed as
`refo_stack_device` for pipelines that stay in HBM.
"""
import numpy as np

from .. import ops
from ..lfp_extractor.lfp_viewpoints import LfpViewpoints


class LfpShiftAndSum(LfpViewpoints):

    #: The reference writes, for every refined slice, the summed UP-SAMPLED image as `upscale_<a/M>.png` through
    #: LfpExporter.save_refo_slice (lfp_shiftandsum.py:126-132; 6510 x 9375 px per slice at Illum size).  File export is
    #: outside this package, so the side effect is opt-in: set `upscale_sink` to a callable `(fname, img)` - e.g.
    #: `lambda a, img: LfpExporter(cfg=cfg, sta=sta).save_refo_slice(a, img, string='upscale_')` with the reference's
    #: exporter - and it receives the same float32 sRGB image under the same name, slice by slice (slow path).
    upscale_sink = None

    def __init__(self, lfp_img_align=None, *args, **kwargs):
        super(LfpShiftAndSum, self).__init__(*args, **kwargs)
        self.lfp_img_align = lfp_img_align
        self.validate_range()
        self._refo_stack = list()
        self._dev_stack = None

    def validate_range(self):
        """lfp_shiftandsum.py:306-315 (may bump cfg.params['ran_refo'][1])."""
        ran = self.cfg.params[self.cfg.ran_refo]
        if len(ran) != 2:
            raise ValueError("Refocusing range list is supposed to contain only two entries")
        if ran[0] >= ran[1]:
            self.cfg.params[self.cfg.ran_refo][1] = ran[0] + 1
            self.sta.status_msg('Refocusing range is not positive')

    def main(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.status_msg('Compute refocused image stack', prnt)
        self.sta.progress(None, prnt)
        if self.vp_img_arr is not None:
            if self.refo_from_vp() is False:
                return False
        elif self.lfp_img_align is not None:
            # the alternates taken when only the aligned image is given (lfp_shiftandsum.py:62-68; `pragma: no cover`
            # upstream, never reached from LfpRefocuser: SURVEY §8a row a11)
            if self.cfg.params[self.cfg.opt_refi]:
                # the reference's own method cannot run: its shape line (:232) makes m the whole shape tuple for a colour
                # image (TypeError at :241) and it indexes a third axis of a monochrome one (IndexError at :258) -
                # tests/test_reference_patch.py::test_reference_scratch_upsample_alternate_cannot_run - so there is no
                # output to reproduce
                raise NotImplementedError('refo_from_scratch_upsample (lfp_shiftandsum.py:223-304: linear ramps between '
                                          'neighbouring micro-image pixels) fails in the reference itself (TypeError / '
                                          'IndexError) and is not offered; pass vp_img_arr for the refinement path')
            if self.refo_from_scratch() is False:
                return False
        return True

    def refo_from_vp(self):
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.progress(0, prnt)
        M = int(self.cfg.params[self.cfg.ptc_leng])
        ran = self.cfg.params[self.cfg.ran_refo]
        vp = self._vp_device()
        if vp.shape[0] != M or vp.shape[1] != M:
            raise ValueError('vp_img_arr angular size %s does not match ptc_leng=%d' % (tuple(vp.shape[:2]), M))
        if vp.shape[-1] != 3:
            # the reference accumulator is hard-wired to 3 channels (lfp_shiftandsum.py:97)
            raise ValueError('operands could not be broadcast together: refocusing needs 3 colour channels')
        if self.cfg.params[self.cfg.opt_refi]:
            from .refinement import refocus_refi
            a_list = range(M * ran[0], M * ran[1])
            self._dev_stack = refocus_refi(vp, a_list, self.circular_view_aperture_mask())
            if self.upscale_sink is not None:
                from .refinement import device_plan, upscaled_slices
                mask = self.circular_view_aperture_mask()
                plan = device_plan(int(vp.shape[2]), int(vp.shape[3]), M, list(a_list), mask, vp.device)
                for a, img in upscaled_slices(vp, list(a_list), mask, plan):
                    if self.sta.interrupt:
                        return False
                    self.upscale_sink(np.round(a / M, 2), np.asarray(ops.to_host(img)))        # fname as in :130
        else:
            a_list = range(*ran)
            self._dev_stack = ops.refocus_int(vp, a_list, self.circular_view_aperture_mask())
        if self.sta.interrupt:
            return False
        self._refo_stack = None              # float64 host copy: made when `refo_stack` is read (LfpRefocuser never does)
        self.sta.progress(100, prnt)
        return True

    def refo_from_scratch(self):
        """lfp_shiftandsum.py:141-221: every view, zeros outside the light field (no aperture mask, no edge replication)."""
        prnt = self.cfg.params[self.cfg.opt_prnt]
        self.sta.progress(0, prnt)
        M = int(self.cfg.params[self.cfg.ptc_leng])
        img = self.lfp_img_align
        on_dev = hasattr(img, 'is_cuda')
        if not on_dev:
            img = np.asarray(img)
            img = ops.to_device(img if img.ndim == 3 else img[..., None], exact_u16=True)
        elif img.dim() == 2:
            img = img[..., None].contiguous()
        rows, cols = int(img.shape[0]), int(img.shape[1])
        img = img[:(rows // M) * M, :(cols // M) * M].contiguous()       # J = int(n / patch_len), H = int(m / patch_len)
        self._dev_stack = ops.refocus_from_scratch(img, M, range(*self.cfg.params[self.cfg.ran_refo]))
        if self.sta.interrupt:
            return False
        self._refo_stack = None
        self.sta.progress(100, prnt)
        return True

    @property
    def refo_stack(self):
        """float64 (N,S,T,C) like the reference's (a fresh array per access, lfp_shiftandsum.py:320-322)."""
        if self._dev_stack is not None:
            return np.asarray(ops.to_host(self._dev_stack), dtype=np.float64)
        return np.asarray(self._refo_stack)

    @property
    def refo_stack_device(self):
        return self._dev_stack
