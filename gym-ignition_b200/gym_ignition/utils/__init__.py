from . import logger, math, misc, resource_finder, typing  # noqa: F401
from . import scenario  # noqa: F401
