from . import gazebo_runtime, realtime_runtime  # noqa: F401
from . import batched_runtime  # noqa: F401
