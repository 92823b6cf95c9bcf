"""Task interface (reference: python/gym_ignition/base/task.py:15-237)."""
import abc
from typing import Dict, Optional, Tuple

import numpy as np
from gym.utils import seeding


class Task(abc.ABC):
    """Decision-making logic of an environment, independent of the runtime that executes it.

    Subclasses implement ``create_spaces / reset_task / set_action / get_observation / get_reward / is_done``
    against the ScenarI/O ``world``. A task may additionally describe itself to the batched engine through
    :py:meth:`batched_spec`, in which case :py:class:`~gym_ignition.runtimes.batched_runtime.BatchedGazeboRuntime`
    evaluates it for all envs inside the fused step kernel instead of calling these methods per env.
    """

    action_space = None
    observation_space = None

    def __init__(self, agent_rate: float) -> None:
        self._world = None
        self.agent_rate = agent_rate
        self.np_random, self.seed = seeding.np_random()

    # -- world handle --
    @property
    def world(self):
        if self._world is None:
            raise Exception("The world was never stored")
        return self._world

    @world.setter
    def world(self, world) -> None:
        if world is None or world.name == "":
            raise ValueError("World not valid")
        self._world = world

    def has_world(self) -> bool:
        return self._world is not None and self._world.name != ""

    # -- interface --
    @abc.abstractmethod
    def create_spaces(self) -> Tuple["gym.spaces.Space", "gym.spaces.Space"]:
        """Return (action_space, observation_space)."""

    @abc.abstractmethod
    def reset_task(self) -> None:
        """Called by ``Env.reset``: put the models in the initial state of a new episode."""

    @abc.abstractmethod
    def set_action(self, action) -> None:
        """Called at the beginning of ``Env.step``."""

    @abc.abstractmethod
    def get_observation(self) -> np.ndarray:
        """Called at the end of ``Env.step`` and ``Env.reset``."""

    @abc.abstractmethod
    def get_reward(self) -> float:
        """Called at the end of ``Env.step``."""

    @abc.abstractmethod
    def is_done(self) -> bool:
        """Called at the end of ``Env.step``."""

    def get_info(self) -> Dict:
        return {}

    def seed_task(self, seed: Optional[int] = None):
        seed = np.random.randint(2 ** 32 - 1) if seed is None else seed
        self.np_random, self.seed = seeding.np_random(seed)
        self.action_space.seed(self.seed)
        self.observation_space.seed(self.seed)
        return [self.seed]

    # -- batched engine hook --
    @classmethod
    def batched_spec(cls) -> Optional[str]:
        """Registered environment id whose fused kernel implements this task, or None."""
        return None
