from .lfp_microlenses import LfpMicroLenses, build_tables  # noqa: F401
from .lfp_resampler import LfpResampler  # noqa: F401
from .lfp_rotator import LfpRotator  # noqa: F401
from .lfp_devignetter import LfpDevignetter  # noqa: F401
from .cfa_processor import CfaProcessor  # noqa: F401
from .top_level import LfpAligner  # noqa: F401
