"""reference: python/gym_ignition_environments/randomizers/cartpole_no_rand.py:29-60."""
from gym_ignition.randomizers import gazebo_env_randomizer
from gym_ignition.randomizers.gazebo_env_randomizer import MakeEnvCallable

from ..models import cartpole


class CartpoleEnvNoRandomizations(gazebo_env_randomizer.GazeboEnvRandomizer):
    """Populates the world for the cart-pole tasks: every reset replaces the cart-pole by a fresh one."""

    def __init__(self, env: MakeEnvCallable):
        super().__init__(env=env)

    def randomize_task(self, task, **kwargs) -> None:
        if "gazebo" not in kwargs:
            raise ValueError("gazebo kwarg not passed to the task randomizer")
        gazebo = kwargs["gazebo"]
        if task.model_name is not None and task.model_name in task.world.model_names():
            if not task.world.to_gazebo().remove_model(task.model_name):
                raise RuntimeError("Failed to remove the cartpole from the world")
        if not gazebo.run(paused=True):  # processes the removal
            raise RuntimeError("Failed to execute a paused Gazebo run")
        task.model_name = cartpole.CartPole(world=task.world).name()
        if not gazebo.run(paused=True):  # processes the insertion
            raise RuntimeError("Failed to execute a paused Gazebo run")
