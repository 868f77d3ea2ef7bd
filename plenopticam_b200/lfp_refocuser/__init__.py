from .lfp_shiftandsum import LfpShiftAndSum  # noqa: F401
from .top_level import LfpRefocuser  # noqa: F401
from .lfp_scheimpflug import LfpScheimpflug  # noqa: F401
