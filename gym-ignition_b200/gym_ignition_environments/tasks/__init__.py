from . import cartpole  # noqa: F401
from . import pendulum_swingup  # noqa: F401
from . import cartpole_discrete_balancing  # noqa: F401
from . import cartpole_continuous_balancing  # noqa: F401
from . import cartpole_continuous_swingup  # noqa: F401
