"""`LfpRefocuser` (lfp_refocuser/top_level.py:32-87): shift-and-sum stack, percentile contrast alignment against
slice 0 and the sRGB transfer, then (opt) the Scheimpflug composites, all on the device.  PNG/GIF export of the
stack (top_level.py:72-77) is out of scope."""
from .. import ops
from ..lfp_extractor.lfp_viewpoints import LfpViewpoints
from .lfp_shiftandsum import LfpShiftAndSum


class LfpRefocuser(LfpViewpoints):

    def __init__(self, vp_img_arr=None, *args, **kwargs):
        super(LfpRefocuser, self).__init__(*args, **kwargs)
        self.vp_img_arr = vp_img_arr
        self.refo_stack = []
        self.refo_stack_device = None

    def main(self):
        if self.sta.interrupt:
            return False
        if self.cfg.params[self.cfg.opt_refo]:
            self.shift_and_sum()
        if self.cfg.params[self.cfg.opt_pflu]:                    # lfp_refocuser/top_level.py:54
            self.scheimpflug()
        return True

    def scheimpflug(self):
        from .lfp_scheimpflug import LfpScheimpflug
        stack = self.refo_stack_device if self.refo_stack_device is not None else self.refo_stack
        obj = LfpScheimpflug(refo_stack=stack, cfg=self.cfg, sta=self.sta)
        obj.main()
        self.scheimpflug_images = obj.images
        del obj
        return True

    def shift_and_sum(self):
        obj = LfpShiftAndSum(vp_img_arr=self.vp_img_arr, cfg=self.cfg, sta=self.sta)
        ok = obj.main()
        dev = obj.refo_stack_device
        del obj
        if ok and dev is not None and not self.sta.interrupt:
            self.refo_stack_device = ops.refocuser_postprocess(dev)
            self.refo_stack = ops.to_host(self.refo_stack_device)      # float32 sRGB, like the reference
        return True
