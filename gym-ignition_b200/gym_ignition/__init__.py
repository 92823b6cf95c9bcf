"""Host-side mirror of the reference's ``gym_ignition`` package for the env.step path.

Same public names and call semantics as /root/reference/python/gym_ignition (Task, Runtime, GazeboRuntime,
utils, randomizers), re-implemented here so the package is self-contained on a machine where the reference
is not installed, plus the batched runtime that drives all envs through one fused kernel launch.
"""
import gym_ignition_models

from . import base, utils  # noqa: F401
from . import scenario  # noqa: F401
from . import randomizers  # noqa: F401
from . import runtimes  # noqa: F401
from . import rbd  # noqa: F401
from . import context  # noqa: F401

gym_ignition_models.setup_environment()
