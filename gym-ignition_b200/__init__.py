"""gym-ignition_b200: B200-native batched engine behind gym-ignition's ScenarI/O surface.

The directory name is not a Python identifier; ``__graft_entry__.load_package()`` (or adding this
directory to ``sys.path``) makes the contained packages importable:

    b2sim                      ctypes binding of the CUDA engine (lib/libb2sim.so, built from csrc/)
    scenario                   drop-in for the reference's SWIG module (scenario.core / scenario.gazebo)
    gym_ignition               host-side mirror of the Task / Runtime interface (+ batched runtime)
    gym_ignition_environments  the registered tasks and model wrappers
    gym_ignition_models        model files
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

try:  # a minimal gym stand-in is used only when the real package is absent
    import gym  # noqa: F401
except ImportError:
    _shims = os.path.join(_HERE, "_shims")
    if os.path.isdir(_shims) and _shims not in sys.path:
        sys.path.append(_shims)

__version__ = "0.1.0"
