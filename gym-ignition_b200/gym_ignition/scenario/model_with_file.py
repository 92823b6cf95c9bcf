"""reference: python/gym_ignition/scenario/model_with_file.py:8-17."""
import abc


class ModelWithFile(abc.ABC):
    def __init__(self):
        super().__init__()

    @classmethod
    @abc.abstractmethod
    def get_model_file(cls) -> str:
        """Path of the URDF / SDF description of the model."""
