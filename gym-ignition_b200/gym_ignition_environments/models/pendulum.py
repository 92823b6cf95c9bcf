"""reference: python/gym_ignition_environments/models/pendulum.py:11-48."""
from typing import List

from gym_ignition.scenario import model_with_file, model_wrapper

from ._insert import insert_named_model


class Pendulum(model_wrapper.ModelWrapper, model_with_file.ModelWithFile):
    def __init__(self, world, position: List[float] = (0.0, 0.0, 0.0), orientation: List[float] = (1.0, 0, 0, 0),
                 model_file: str = None):
        super().__init__(model=insert_named_model(world, "pendulum", position, orientation, model_file))

    @classmethod
    def get_model_file(cls) -> str:
        import gym_ignition_models
        return gym_ignition_models.get_model_file("pendulum")
