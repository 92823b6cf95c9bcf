"""Interface of the model wrappers that know where their description lives.

Mirrors the role of ``ModelWithFile`` in the reference (python/gym_ignition/scenario/model_with_file.py:8-17): the
wrappers of ``gym_ignition_environments.models`` (cartpole, pendulum, panda) and the randomizers ask the CLASS, not an
instance, for the URDF / SDF file to insert into a world.
"""
from abc import ABC, abstractmethod


class ModelWithFile(ABC):
    """Mixin: ``cls.get_model_file()`` returns the path handed to ``World.insert_model``."""

    @classmethod
    @abstractmethod
    def get_model_file(cls) -> str:
        raise NotImplementedError("model wrappers return the path of their URDF / SDF file here")
