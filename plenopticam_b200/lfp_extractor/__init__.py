from .lfp_viewpoints import LfpViewpoints  # noqa: F401
from .lfp_cropper import LfpCropper  # noqa: F401
from .lfp_rearranger import LfpRearranger  # noqa: F401
from .lfp_contrast import LfpContrast  # noqa: F401
from .top_level import LfpExtractor  # noqa: F401
