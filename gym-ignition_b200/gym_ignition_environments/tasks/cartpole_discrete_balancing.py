"""reference: python/gym_ignition_environments/tasks/cartpole_discrete_balancing.py:41-144."""
from .cartpole import CartPoleBalancingTask


class CartPoleDiscreteBalancing(CartPoleBalancingTask):
    """Discrete(2) action -> +-20 N on the cart; rail-end penalty from 0.9 * 2.4 m."""
    max_force = None
    force_mag = 20.0
    edge_fraction = 0.9
    env_id = "CartPoleDiscreteBalancing-Gazebo-v0"

    def __init__(self, agent_rate: float, reward_cart_at_center: bool = True, **kwargs) -> None:
        super().__init__(agent_rate, reward_cart_at_center, **kwargs)
        self._force_mag = self.force_mag
