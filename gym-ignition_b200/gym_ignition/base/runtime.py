"""Runtime interface (reference: python/gym_ignition/base/runtime.py:10-81)."""
import abc

import gym


class Runtime(gym.Env, abc.ABC):
    """Executor of a Task: the ``gym.Env`` users get from ``gym.make``. One runtime handles one task."""

    def __init__(self, task, agent_rate: float):
        self.task = task
        self.agent_rate = agent_rate

    @abc.abstractmethod
    def timestamp(self) -> float:
        """Time associated with the execution of the environment (simulated time for simulators)."""
