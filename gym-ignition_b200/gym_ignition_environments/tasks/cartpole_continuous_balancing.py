"""reference: python/gym_ignition_environments/tasks/cartpole_continuous_balancing.py:40-146."""
from .cartpole import CartPoleBalancingTask


class CartPoleContinuousBalancing(CartPoleBalancingTask):
    """Box(+-50 N) action; rail-end penalty at 2.4 m."""
    max_force = 50.0
    edge_fraction = 1.0
    env_id = "CartPoleContinuousBalancing-Gazebo-v0"
