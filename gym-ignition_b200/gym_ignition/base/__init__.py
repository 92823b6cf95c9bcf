from . import task, runtime  # noqa: F401
