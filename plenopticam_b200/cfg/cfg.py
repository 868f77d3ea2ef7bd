"""Stand-in for the reference's config object so the drop-in stages run without the reference installed.

Only the surface the hot path touches is mirrored (cfg/cfg.py:37-66,130-154,230-244 and
cfg/constants.py:2-63,128-132): the `params` / `calibs` / `lfpimg` dicts, the key-name attributes and
`exp_path`.  The real `plenopticam.cfg.PlenopticamConfig` is accepted everywhere instead (duck typing);
nothing here writes a cfg.json.
"""
import os

PARAMS_KEYS = ('lfp_path', 'cal_path', 'cal_meta', 'cal_meth', 'smp_meth', 'ptc_leng', 'ran_refo',
               'opt_cali', 'opt_vign', 'opt_lier', 'opt_cont', 'opt_colo', 'opt_awb_', 'opt_sat_', 'opt_view',
               'opt_refo', 'opt_refi', 'opt_pflu', 'opt_arti', 'opt_rota', 'opt_dbug', 'opt_prnt', 'opt_dpth',
               'dir_remo')
PARAMS_VALS = ('', '', '', '', '', 7, [0, 2],
               False, True, False, False, True, False, False, True,
               True, False, False, False, False, False, True, True,
               False)
CALIBS_KEYS = ('pat_type', 'ptc_mean', 'mic_list')
SMPL_METH = ('global', 'local')


class PlenopticamConfig(object):

    def __init__(self):
        self.params = {}
        self.calibs = {}
        self.lfpimg = {}
        self.default_values()

    def default_values(self):
        self.params = {k: (list(v) if isinstance(v, list) else v) for k, v in zip(PARAMS_KEYS, PARAMS_VALS)}
        return True

    @property
    def exp_path(self):
        p = self.params['lfp_path']
        return os.path.splitext(p)[0] if p else os.path.expanduser("~")

    def cond_lfp_align(self):
        return not os.path.exists(os.path.join(self.exp_path, 'lfp_img_align.pkl'))


for _k in PARAMS_KEYS + CALIBS_KEYS:
    setattr(PlenopticamConfig, _k, _k)
