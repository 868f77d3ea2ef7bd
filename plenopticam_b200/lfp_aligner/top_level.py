"""`LfpAligner` restricted to the hot path: (opt) LfpRotator -> LfpResampler
(lfp_aligner/top_level.py:32-83).  The pre-processing stages that precede them in the reference
(CfaOutliers, LfpDevignetter, CfaProcessor; top_level.py:44-64) are out of scope — feed this
class the RGB image those stages produce.  The rotated image stays on the device between the two stages."""
from ..misc import PlenopticamStatus
from .lfp_resampler import LfpResampler
from .lfp_rotator import LfpRotator


class LfpAligner(object):

    def __init__(self, lfp_img, cfg=None, sta=None, wht_img=None):
        self.cfg = cfg
        self.sta = sta if sta is not None else PlenopticamStatus()
        self._lfp_img = lfp_img
        self._wht_img = wht_img
        if cfg is not None and getattr(cfg, 'lfpimg', None):
            raise NotImplementedError('Bayer/Lytro pre-processing (CfaOutliers, CfaProcessor) is outside the B200 '
                                      'hot path; pass the demosaiced RGB image')

    def main(self):
        img = self._lfp_img
        if self.cfg.params[self.cfg.opt_rota] and img is not None:
            obj = LfpRotator(img, self.cfg.calibs[self.cfg.mic_list], rad=None, cfg=self.cfg, sta=self.sta)
            obj.main()
            img = obj.lfp_img_device if obj.lfp_img_device is not None else img
            self.cfg.calibs[self.cfg.mic_list] = obj.centroids
            del obj
        obj = LfpResampler(lfp_img=img, cfg=self.cfg, sta=self.sta)
        obj.main()
        self._lfp_img = obj.lfp_img_align
        del obj
        return True

    @property
    def lfp_img(self):
        return self._lfp_img.copy()
