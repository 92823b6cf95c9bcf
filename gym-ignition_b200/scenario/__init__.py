"""Drop-in for the reference's SWIG-generated ``scenario`` package (bindings/__init__.py:134-147):

    from scenario import core
    from scenario import gazebo as scenario

backed by the B200 engine (``b2sim``) instead of Ignition Gazebo + DART.
"""
from . import core
from . import gazebo
from . import bindings

__all__ = ["core", "gazebo", "bindings"]
