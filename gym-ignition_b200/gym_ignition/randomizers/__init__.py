from . import abc  # noqa: F401
from . import physics  # noqa: F401
from . import gazebo_env_randomizer  # noqa: F401
