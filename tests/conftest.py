import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from plenopticam_b200 import _native
    _native.lib()          # must load: no fallback
    return torch.device("cuda", 0)


def make_cfg(mic_list=None, pat_type='rec', M=7, ran_refo=(0, 2), refi=False, tmpdir=None):
    from plenopticam_b200.cfg import PlenopticamConfig
    from plenopticam_b200.misc import PlenopticamStatus
    cfg = PlenopticamConfig()
    cfg.params[cfg.smp_meth] = 'local'
    cfg.params[cfg.ptc_leng] = int(M)
    cfg.params[cfg.ran_refo] = list(ran_refo)
    cfg.params[cfg.opt_refi] = bool(refi)
    cfg.params[cfg.opt_prnt] = False
    if tmpdir is not None:
        cfg.params[cfg.lfp_path] = os.path.join(str(tmpdir), "exposure.png")
    if mic_list is not None:
        cfg.calibs[cfg.mic_list] = np.asarray(mic_list).tolist()
        cfg.calibs[cfg.pat_type] = pat_type
        cfg.calibs[cfg.ptc_mean] = [float(M), float(M)]
    return cfg, PlenopticamStatus()
