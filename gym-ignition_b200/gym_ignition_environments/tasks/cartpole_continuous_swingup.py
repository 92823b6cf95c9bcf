"""reference: python/gym_ignition_environments/tasks/cartpole_continuous_swingup.py:40-153."""
import numpy as np

from .cartpole import CartPoleTask


class CartPoleContinuousSwingup(CartPoleTask):
    """Box(+-200 N) action, pole starting near the bottom, reward (cos q + 1)/2 - 0.1 dx^2 - 10 [x >= 0.8*2.4]."""
    max_force = 200.0
    q_threshold = np.deg2rad(5 * 360)
    env_id = "CartPoleContinuousSwingup-Gazebo-v0"

    def get_reward(self) -> float:
        model = self._model()
        q = model.get_joint("pivot").position()
        x = model.get_joint("linear").position()
        dx = model.get_joint("linear").velocity()
        reward = (np.cos(q) + 1) / 2
        reward -= 0.1 * (dx ** 2)
        reward -= 10.0 * (x >= 0.8 * self._x_threshold)
        return reward

    def _sample_state(self):
        # draw order matters: q first, then x, dx, dq (cartpole_continuous_swingup.py:145-146)
        q = np.pi - np.deg2rad(self.np_random.uniform(low=-60, high=60))
        x, dx, dq = self.np_random.uniform(low=-0.05, high=0.05, size=(3,))
        return x, dx, q, dq
