"""`LfpExtractor` restricted to the hot path: LfpCropper -> LfpRearranger
(lfp_extractor/top_level.py:53-73).  Colour equalisation, contrast, hex-artefact removal, export and depth
(top_level.py:75-118) are out of scope; `vp_img_linear` is therefore exactly the gathered array,
which is what the reference hands to LfpRefocuser when those options are off (SURVEY §8c ii)."""
from ..misc import PlenopticamStatus
from .lfp_cropper import LfpCropper
from .lfp_rearranger import LfpRearranger


class LfpExtractor(object):

    def __init__(self, lfp_img_align=None, cfg=None, sta=None):
        self._lfp_img_align = lfp_img_align
        self.cfg = cfg
        self.sta = sta if sta is not None else PlenopticamStatus()
        self.vp_img_arr = None
        self.vp_img_linear = None

    def main(self):
        if self.sta.interrupt:
            return False
        obj = LfpCropper(lfp_img_align=self._lfp_img_align, cfg=self.cfg, sta=self.sta)
        obj.main()
        self._lfp_img_align = obj.lfp_img_align
        del obj
        if self.cfg.params[self.cfg.opt_view] or self.cfg.params[self.cfg.opt_refo]:
            obj = LfpRearranger(self._lfp_img_align, cfg=self.cfg, sta=self.sta)
            obj.main()
            self.vp_img_linear = obj.vp_img_arr
            self.vp_img_arr = self.vp_img_linear
            del obj
        return True
