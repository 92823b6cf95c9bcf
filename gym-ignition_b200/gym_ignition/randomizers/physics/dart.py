"""reference: python/gym_ignition/randomizers/physics/dart.py."""
from scenario import gazebo as scenario

from .. import abc as randomizers_abc


class DART(randomizers_abc.PhysicsRandomizer):
    """DART-equivalent physics without randomization (the default of every runtime)."""

    def __init__(self):
        super().__init__()

    def get_engine(self):
        return scenario.PhysicsEngine_dart

    def randomize_physics(self, task, **kwargs) -> None:
        if not task.world.to_gazebo().set_physics_engine(scenario.PhysicsEngine_dart):
            raise RuntimeError("Failed to insert the physics plugin")
