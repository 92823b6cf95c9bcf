"""reference: python/gym_ignition_environments/tasks/pendulum_swingup.py:29-130."""
from typing import Tuple

import gym
import numpy as np
from gym_ignition.base import task
from scenario import core as scenario


class PendulumSwingUp(task.Task):
    """Torque on ``pivot`` in [-50, 50] N m, observation [cos q, sin q, dq], done when |dq| > 10."""
    env_id = "Pendulum-Gazebo-v0"

    def __init__(self, agent_rate: float, **kwargs):
        task.Task.__init__(self, agent_rate=agent_rate)
        self.model_name = None
        self._max_speed = 10.0
        self._max_torque = 50.0

    @classmethod
    def batched_spec(cls):
        return cls.env_id

    def create_spaces(self) -> Tuple[gym.spaces.Space, gym.spaces.Space]:
        action_space = gym.spaces.Box(low=-self._max_torque, high=self._max_torque, shape=(1,), dtype=np.float32)
        high = np.array([1.0, 1.0, self._max_speed])
        return action_space, gym.spaces.Box(low=-high, high=high, dtype=np.float32)

    def _pivot(self):
        return self.world.get_model(self.model_name).get_joint("pivot")

    def set_action(self, action) -> None:
        if not self._pivot().set_generalized_force_target(action.tolist()[0]):
            raise RuntimeError("Failed to set the force to the pendulum")

    def get_observation(self) -> np.ndarray:
        pivot = self._pivot()
        q, dq = pivot.position(), pivot.velocity()
        return np.array([np.cos(q), np.sin(q), dq])

    def is_done(self) -> bool:
        return not self.observation_space.contains(self.get_observation())

    def get_reward(self) -> float:
        cost = 100.0 if self.is_done() else 0.0
        pivot = self._pivot()
        q, dq = pivot.position(), pivot.velocity()
        tau = pivot.generalized_force_target()  # already zeroed by the physics step (Physics.cpp:2250-2254)
        cost += (q ** 2) + 0.1 * (dq ** 2) + 0.001 * (tau ** 2)
        return float(-cost)

    def reset_task(self) -> None:
        if self.model_name not in self.world.model_names():
            raise RuntimeError("The pendulum model was not inserted in the world")
        pivot = self._pivot()
        if not pivot.set_control_mode(scenario.JointControlMode_force):
            raise RuntimeError("Failed to change the control mode of the pendulum")
        cos_q, sin_q, dq = self.observation_space.sample()
        q = np.arctan2(sin_q, cos_q)
        if not pivot.to_gazebo().reset(float(q), float(dq)):
            raise RuntimeError("Failed to reset the pendulum state")
