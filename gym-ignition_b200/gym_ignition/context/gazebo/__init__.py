from . import plugin  # noqa: F401
from . import controllers  # noqa: F401
