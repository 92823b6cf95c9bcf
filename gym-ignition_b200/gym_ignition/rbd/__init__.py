from . import conversions, utils  # noqa: F401
from . import kindyn  # noqa: F401
