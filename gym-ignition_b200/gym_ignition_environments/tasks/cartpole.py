"""Shared implementation of the three cart-pole tasks.

The reference has one file per task (python/gym_ignition_environments/tasks/cartpole_*.py) that differ only in
the action space, the termination bounds, the reward and the reset distribution; here those are class
attributes of one base. Observation is [x, dx, q, dq] read through ScenarI/O in the joint order
["pivot", "linear"]; termination is ``not reset_space.contains(observation)`` on float32-rounded bounds.
"""
import abc
from typing import Tuple

import gym
import numpy as np
from gym_ignition.base import task
from scenario import core as scenario


class CartPoleTask(task.Task, abc.ABC):
    #: force limit of the continuous action space [N]; None = Discrete(2) with +-force_mag
    max_force = None
    force_mag = 20.0
    #: termination bounds [x, dx, q, dq]
    x_threshold = 2.4
    dx_threshold = 20.0
    q_threshold = np.deg2rad(12)
    dq_threshold = np.deg2rad(3 * 360)
    #: registered id of the fused kernel that implements the task
    env_id = None

    def __init__(self, agent_rate: float, reward_cart_at_center: bool = True, **kwargs):
        task.Task.__init__(self, agent_rate=agent_rate)
        self.model_name = None
        self.reset_space = None
        self._reward_cart_at_center = reward_cart_at_center
        self._x_threshold, self._dx_threshold = self.x_threshold, self.dx_threshold
        self._q_threshold, self._dq_threshold = self.q_threshold, self.dq_threshold

    @classmethod
    def batched_spec(cls):
        return cls.env_id

    # -- spaces --
    def create_spaces(self) -> Tuple[gym.spaces.Space, gym.spaces.Space]:
        if self.max_force is None:
            action_space = gym.spaces.Discrete(2)
        else:
            action_space = gym.spaces.Box(low=np.array([-self.max_force]), high=np.array([self.max_force]),
                                          dtype=np.float32)
        high = np.array([self._x_threshold, self._dx_threshold, self._q_threshold, self._dq_threshold])
        self.reset_space = gym.spaces.Box(low=-high, high=high, dtype=np.float32)
        obs_high = high.copy() * 1.2
        return action_space, gym.spaces.Box(low=-obs_high, high=obs_high, dtype=np.float32)

    # -- step --
    def _model(self):
        return self.world.get_model(self.model_name)

    def _force_of(self, action) -> float:
        if self.max_force is None:
            return self.force_mag if action == 1 else -self.force_mag
        return action.tolist()[0]

    def set_action(self, action) -> None:
        if not self._model().get_joint("linear").set_generalized_force_target(self._force_of(action)):
            raise RuntimeError("Failed to set the force to the cart")

    def get_observation(self) -> np.ndarray:
        model = self._model()
        q, x = model.joint_positions(["pivot", "linear"])
        dq, dx = model.joint_velocities(["pivot", "linear"])
        return np.array([x, dx, q, dq])

    def is_done(self) -> bool:
        return not self.reset_space.contains(self.get_observation())

    # -- reset --
    @abc.abstractmethod
    def _sample_state(self) -> Tuple[float, float, float, float]:
        """Return x, dx, q, dq of a fresh episode (draw order is part of the task definition)."""

    def reset_task(self) -> None:
        if self.model_name not in self.world.model_names():
            raise RuntimeError("Cartpole model not found in the world")
        model = self._model()
        if not model.get_joint("linear").set_control_mode(scenario.JointControlMode_force):
            raise RuntimeError("Failed to change the control mode of the cartpole")
        x, dx, q, dq = self._sample_state()
        gazebo_model = model.to_gazebo()
        ok_pos = gazebo_model.reset_joint_positions([x, q], ["linear", "pivot"])
        ok_vel = gazebo_model.reset_joint_velocities([dx, dq], ["linear", "pivot"])
        if not (ok_pos and ok_vel):
            raise RuntimeError("Failed to reset the cartpole state")


class CartPoleBalancingTask(CartPoleTask, abc.ABC):
    """reward = alive bonus - 0.1 |x| - 0.1 |dx| - 10 [x >= edge] (cartpole_discrete_balancing.py:94-109)."""

    #: fraction of x_threshold where the rail-end penalty starts
    edge_fraction = 1.0

    def get_reward(self) -> float:
        reward = 1.0 if not self.is_done() else 0.0
        if self._reward_cart_at_center:
            x, dx, _, _ = self.get_observation()
            reward = reward - 0.10 * np.abs(x) - 0.10 * np.abs(dx) - 10.0 * (x >= self._edge())
        return reward

    def _edge(self) -> float:
        return self._x_threshold if self.edge_fraction == 1.0 else self.edge_fraction * self._x_threshold

    def _sample_state(self):
        x, dx, q, dq = self.np_random.uniform(low=-0.05, high=0.05, size=(4,))
        return x, dx, q, dq
