"""Environment randomizer wrapper (reference: python/gym_ignition/randomizers/gazebo_env_randomizer.py:16-142)."""
import abc
from typing import Callable, Dict, Optional, Union

import gym

from ..runtimes import gazebo_runtime
from ..utils import logger
from . import abc as randomizers_abc
from .physics import dart

MakeEnvCallable = Callable[[Optional[Dict]], gym.Env]


class GazeboEnvRandomizer(gym.Wrapper, randomizers_abc.TaskRandomizer, abc.ABC):
    """``gym.Wrapper`` whose ``reset`` (re)populates the world through ``randomize_task`` before the task resets.

    ``env`` is either a registered id or a callable returning the environment. When the physics randomizer
    reports that physics expired, the whole runtime is closed and created again (a new simulator).
    """

    def __init__(self, env: Union[str, MakeEnvCallable],
                 physics_randomizer: randomizers_abc.PhysicsRandomizer = None, **kwargs):
        physics_randomizer = physics_randomizer if physics_randomizer is not None else dart.DART()
        self._env_option = env
        self._kwargs = dict(**kwargs, physics_engine=physics_randomizer.get_engine())
        self._physics_randomizer = physics_randomizer
        gym.Wrapper.__init__(self, env=self._make())

    def _make(self):
        with logger.gym_verbosity(level=gym.logger.WARN):
            if isinstance(self._env_option, str):
                env = gym.make(self._env_option, **self._kwargs)
            elif callable(self._env_option):
                env = self._env_option(**self._kwargs)
            else:
                raise ValueError("The type of env object was not recognized")
        if not isinstance(env.unwrapped, gazebo_runtime.GazeboRuntime):
            raise ValueError("The environment to wrap is not a GazeboRuntime")
        return env

    def reset(self, **kwargs):
        if self._physics_randomizer.physics_expired():
            seed, rng = self.env.task.seed, self.env.task.np_random
            self.env.close()
            del self.env
            self.env = self._make()
            self.env.seed(seed=seed)
            assert self.env.task.seed == seed
            self.env.task.np_random = rng
        self._physics_randomizer.increase_rollout_counter()
        self.randomize_task(task=self.env.task, gazebo=self.env.gazebo, **kwargs)
        if not self.env.gazebo.run(paused=True):
            raise RuntimeError("Failed to execute a paused Gazebo run")
        return self.env.reset()
