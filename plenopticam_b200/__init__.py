"""plenopticam_b200 — B200-native (sm_100a) drop-in for PlenoptiCam's data-parallel light-field hot path:
LfpResampler / LfpRotator (align), LfpCropper / LfpRearranger (viewpoints), LfpShiftAndSum / LfpRefocuser
(refocus).  Host classes mirror the reference's; the arithmetic runs in hand-written CUDA behind a C-ABI
(include/plenopticam_b200.h).  There is no CPU fallback."""
__version__ = "0.1.0"

from . import cfg, misc  # noqa: F401
from .cfg import PlenopticamConfig  # noqa: F401
from .misc import PlenopticamStatus, PlenopticamError  # noqa: F401
