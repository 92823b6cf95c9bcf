from .space import Space
from .box import Box
from .discrete import Discrete

__all__ = ["Space", "Box", "Discrete"]
