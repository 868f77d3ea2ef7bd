from .status import PlenopticamStatus  # noqa: F401
from .errors import PlenopticamError  # noqa: F401
from .type_checks import rint  # noqa: F401
