from . import gazebo  # noqa: F401
