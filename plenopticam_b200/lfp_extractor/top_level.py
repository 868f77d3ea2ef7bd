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
        self._vp_img_arr = None
        self._vp_srgb_dev = None          # sRGB float32 views on the device; copied to the host when `vp_img_arr` is read
        self.vp_img_linear = None
        self.depth_map = None
        self.skipped = []

    @property
    def vp_img_arr(self):
        """Histogram-aligned, sRGB-encoded float32 views (lfp_extractor/top_level.py:87-94).  For host input the 4-D array
        (732 MB at Illum size) crosses PCIe when it is first read, not inside main()."""
        if self._vp_img_arr is None and self._vp_srgb_dev is not None:
            from .. import ops
            self._vp_img_arr = ops.to_host(self._vp_srgb_dev)
            self._vp_srgb_dev = None
        return self._vp_img_arr

    @vp_img_arr.setter
    def vp_img_arr(self, value):
        self._vp_img_arr = value
        self._vp_srgb_dev = None

    def main(self):
        if self.sta.interrupt:
            return False
        prnt = self.cfg.params.get(getattr(self.cfg, 'opt_prnt', 'opt_prnt'), False)
        dev_vp = None
        obj = LfpCropper(lfp_img_align=self._lfp_img_align, cfg=self.cfg, sta=self.sta)
        obj.main()
        self._lfp_img_align = obj.lfp_img_align
        del obj
        if self.cfg.params[self.cfg.opt_view]:                     # lfp_extractor/top_level.py:68
            obj = LfpRearranger(self._lfp_img_align, cfg=self.cfg, sta=self.sta)
            obj.main()
            self.vp_img_linear = obj.vp_img_arr
            dev_vp = getattr(obj, '_dev_vp', None)
            del obj
        for opt, stage in _OUT_OF_SCOPE:
            if self.cfg.params.get(getattr(self.cfg, opt, opt), False):
                self.skipped.append(stage)
                self.sta.status_msg('%s skipped: outside the B200 path' % stage, prnt)
        self.vp_img_arr = self.vp_img_linear
        if self.vp_img_linear is not None:
            on_dev = hasattr(self.vp_img_linear, 'is_cuda')
            # the gather's device result feeds the colour management directly (no second upload of the light field)
            src = dev_vp if dev_vp is not None else self.vp_img_linear
            obj = LfpContrast(vp_img_arr=src, cfg=self.cfg, sta=self.sta)
            obj.main()
            if on_dev:
                self.vp_img_arr = obj.vp_img_arr_device
            else:
                self._vp_img_arr, self._vp_srgb_dev = None, obj.vp_img_arr_device
            del obj
        return True
