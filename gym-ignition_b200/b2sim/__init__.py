"""b2sim: Python binding of the B200-native batched physics-and-observation engine."""
from . import _lib
from ._lib import B2Error
from .engine import DeviceArray, ModelInfo, Simulator

__all__ = ["_lib", "B2Error", "DeviceArray", "ModelInfo", "Simulator"]
from .batched import BatchedTaskEnv  # noqa: E402

__all__.append("BatchedTaskEnv")
from .scenes import PandaPickScene  # noqa: E402

__all__.append("PandaPickScene")
