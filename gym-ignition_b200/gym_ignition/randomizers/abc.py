"""Randomizer interfaces (reference: python/gym_ignition/randomizers/abc.py:10-134)."""
import abc


class TaskRandomizer(abc.ABC):
    @abc.abstractmethod
    def randomize_task(self, task, **kwargs) -> None:
        """Prepare / randomize the world of ``task`` for a new rollout."""


class PhysicsRandomizer(abc.ABC):
    """Decides when the physics must be re-created (``randomize_after_rollouts_num`` resets; 0 = never)."""

    def __init__(self, randomize_after_rollouts_num: int = 0):
        self.randomize_after_rollouts_num = randomize_after_rollouts_num
        self._rollout_counter = randomize_after_rollouts_num

    @abc.abstractmethod
    def randomize_physics(self, task, **kwargs) -> None:
        """Insert and configure the physics of the task's world."""

    @abc.abstractmethod
    def get_engine(self):
        """Physics engine enum to load."""

    def increase_rollout_counter(self) -> None:
        if self.randomize_after_rollouts_num != 0:
            assert self._rollout_counter != 0
            self._rollout_counter -= 1

    def physics_expired(self) -> bool:
        if self.randomize_after_rollouts_num == 0:
            return False
        if self._rollout_counter == 0:
            self._rollout_counter = self.randomize_after_rollouts_num
            return True
        return False


class ModelRandomizer(abc.ABC):
    @abc.abstractmethod
    def randomize_model(self, task, **kwargs):
        """Return the randomized model."""


class ModelDescriptionRandomizer(abc.ABC):
    @abc.abstractmethod
    def randomize_model_description(self, task, **kwargs) -> str:
        """Return the randomized model description (URDF / SDF string)."""
