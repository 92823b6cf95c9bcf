"""``scenario.bindings.{core,gazebo}`` aliases, for code that imports the SWIG module paths directly."""
from .. import core, gazebo

__all__ = ["core", "gazebo"]
