"""`LfpExtractor` restricted to the stages that are part of the B200 path: LfpCropper -> LfpRearranger ->
LfpContrast (lfp_extractor/top_level.py:53-97).  `vp_img_linear` is the gathered light field (what the reference hands
to LfpRefocuser), `vp_img_arr` its histogram-aligned, sRGB-encoded float32 copy.

Outlier removal, colour equalisation (third-party color_matcher), hex-artefact removal, export and depth
(third-party depthy) are out of scope: when their options are set the stage is skipped, reported through
`sta.status_msg` and listed in `skipped` — with those options off the results equal the reference's."""
from ..misc import PlenopticamStatus
from .lfp_contrast import LfpContrast
from .lfp_cropper import LfpCropper
from .lfp_rearranger import LfpRearranger

_OUT_OF_SCOPE = (('opt_lier', 'LfpOutliers'), ('opt_colo', 'LfpColorEqualizer'), ('opt_arti', 'HexCorrector'),
                 ('opt_dpth', 'LfpDepth'))


class LfpExtractor(object):

    def __init__(self, lfp_img_align=None, cfg=None, sta=None):
        self._lfp_img_align = lfp_img_align
        self.cfg = cfg
        self.sta = sta if sta is not None else PlenopticamStatus()
        self.vp_img_arr = None
        self.vp_img_linear = None
        self.depth_map = None
        self.skipped = []

    def main(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params.get(getattr(self.cfg, 'opt_prnt', 'opt_prnt'), False)
        obj = LfpCropper(lfp_img_align=self._lfp_img_align, cfg=self.cfg, sta=self.sta)
        obj.main()
        self._lfp_img_align = obj.lfp_img_align
        del obj
        if self.cfg.params[self.cfg.opt_view]:                     # lfp_extractor/top_level.py:68
            obj = LfpRearranger(self._lfp_img_align, cfg=self.cfg, sta=self.sta)
            obj.main()
            self.vp_img_linear = obj.vp_img_arr
            del obj
        for opt, stage in _OUT_OF_SCOPE:
            if self.cfg.params.get(getattr(self.cfg, opt, opt), False):
                self.skipped.append(stage)
                self.sta.status_msg('%s skipped: outside the B200 path' % stage, prnt)
        self.vp_img_arr = self.vp_img_linear
        if self.vp_img_linear is not None:
            obj = LfpContrast(vp_img_arr=self.vp_img_linear, cfg=self.cfg, sta=self.sta)
            obj.main()
            on_dev = hasattr(self.vp_img_linear, 'is_cuda')
            self.vp_img_arr = obj.vp_img_arr_device if on_dev else obj.vp_img_arr
            del obj
        return True
