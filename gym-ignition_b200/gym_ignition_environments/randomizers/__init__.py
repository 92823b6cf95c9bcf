from . import cartpole_no_rand  # noqa: F401
