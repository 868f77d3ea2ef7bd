"""`LfpAligner` for RGB exposures: (opt) LfpDevignetter -> (opt) LfpRotator -> LfpResampler
(lfp_aligner/top_level.py:32-83).  The Bayer stages of Lytro files (CfaOutliers, CfaProcessor; top_level.py:44-49,
59-64) are out of scope — feed this class the RGB image they produce.  The image stays on the device
between the stages."""
from ..misc import PlenopticamStatus
from .lfp_devignetter import LfpDevignetter
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
        if self.cfg.params[self.cfg.opt_vign] and self._wht_img is not None and img is not None:
            obj = LfpDevignetter(lfp_img=img, wht_img=self._wht_img, cfg=self.cfg, sta=self.sta)
            obj.main()
            img = obj.lfp_img_device
            self._wht_dev = obj._dev_wn
            del obj
        if self.cfg.params[self.cfg.opt_rota] and img is not None:
            obj = LfpRotator(img, self.cfg.calibs[self.cfg.mic_list], rad=None, cfg=self.cfg, sta=self.sta)
            obj.main()
            img = obj.lfp_img_device if obj.lfp_img_device is not None else img
            self.cfg.calibs[self.cfg.mic_list] = obj.centroids
            del obj
        obj = LfpResampler(lfp_img=img, cfg=self.cfg, sta=self.sta)
        obj.main()
        self._lfp_img = obj._lfp_img_align             # the result array itself (the property would copy 366 MB)
        del obj
        return True

    @property
    def lfp_img(self):
        """The aligned uint16 image.  The reference returns a defensive copy (lfp_aligner/top_level.py:81-83); here the
        array itself is handed over (it lives in pinned memory, so LfpExtractor uploads it at the full PCIe rate) - this
        object does not touch it again."""
        return self._lfp_img
