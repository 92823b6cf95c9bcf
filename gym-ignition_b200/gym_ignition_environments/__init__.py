"""Registered environments (reference: python/gym_ignition_environments/__init__.py:12-52): same ids, entry
point, rates (agent 1000 Hz, physics 1000 Hz, uncapped real-time factor) and 5000-step episode cap."""
import numpy
from gym.envs.registration import register

from . import models, randomizers, tasks  # noqa: F401

max_float = float(numpy.finfo(numpy.float32).max)

_COMMON = {"agent_rate": 1000, "physics_rate": 1000, "real_time_factor": max_float}

for _env_id, _task_cls in (
        ("Pendulum-Gazebo-v0", tasks.pendulum_swingup.PendulumSwingUp),
        ("CartPoleDiscreteBalancing-Gazebo-v0", tasks.cartpole_discrete_balancing.CartPoleDiscreteBalancing),
        ("CartPoleContinuousBalancing-Gazebo-v0", tasks.cartpole_continuous_balancing.CartPoleContinuousBalancing),
        ("CartPoleContinuousSwingup-Gazebo-v0", tasks.cartpole_continuous_swingup.CartPoleContinuousSwingup)):
    register(id=_env_id, entry_point="gym_ignition.runtimes.gazebo_runtime:GazeboRuntime", max_episode_steps=5000,
             kwargs=dict(task_cls=_task_cls, **_COMMON))
