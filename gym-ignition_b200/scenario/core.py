"""``scenario.core``: the engine-agnostic half of ScenarI/O, as the reference's SWIG module exposes it.

Names follow bindings/core/core.i: C++ camelCase methods become snake_case (core.i:70), ``enum class E{Foo}``
becomes the module-level int ``E_foo``, every C++ exception becomes ``RuntimeError`` (core.i:14-22).
Types mirror cpp/scenario/core/include/scenario/core/{Joint,Link,Model,World}.h.
"""
import abc
import sys
from typing import List, Sequence

_DBL_MAX = sys.float_info.max

# enum class JointType (core/Joint.h:25-33)
JointType_invalid, JointType_fixed, JointType_revolute, JointType_prismatic, JointType_ball = range(5)

# enum class JointControlMode (core/Joint.h:35-45)
(JointControlMode_invalid, JointControlMode_idle, JointControlMode_force, JointControlMode_velocity,
 JointControlMode_velocity_follower_dart, JointControlMode_position,
 JointControlMode_position_interpolated) = range(7)


class PID:
    """core/Joint.h:505-523. ``PID(p, i, d)`` leaves the integral and command limits at +-DBL_MAX."""

    def __init__(self, p: float = 0.0, i: float = 0.0, d: float = 0.0):
        self.p, self.i, self.d = float(p), float(i), float(d)
        self.cmd_min, self.cmd_max = -_DBL_MAX, _DBL_MAX
        self.cmd_offset = 0.0
        self.i_min, self.i_max = -_DBL_MAX, _DBL_MAX

    def __repr__(self):
        return (f"PID(p={self.p}, i={self.i}, d={self.d}, i_min={self.i_min}, i_max={self.i_max}, "
                f"cmd_min={self.cmd_min}, cmd_max={self.cmd_max}, cmd_offset={self.cmd_offset})")


class Limit:
    """core/Joint.h:525-535."""

    def __init__(self, min: float = -_DBL_MAX, max: float = _DBL_MAX):
        self.min, self.max = float(min), float(max)


class JointLimit:
    """core/Joint.h:537-563: per-DoF vectors of lower / upper bounds."""

    def __init__(self, *args):
        if len(args) == 2:
            lo, hi = list(args[0]), list(args[1])
            if len(lo) != len(hi):
                raise RuntimeError("The max and min limits have different size")
            self.min, self.max = tuple(float(v) for v in lo), tuple(float(v) for v in hi)
        else:
            dofs = int(args[0]) if args else 0
            self.min, self.max = tuple([-_DBL_MAX] * dofs), tuple([_DBL_MAX] * dofs)


class Pose:
    """core/Link.h / World.h ``Pose``: position xyz and orientation quaternion wxyz."""

    def __init__(self, position: Sequence[float] = (0.0, 0.0, 0.0),
                 orientation: Sequence[float] = (1.0, 0.0, 0.0, 0.0)):
        self.position = tuple(float(v) for v in position)
        self.orientation = tuple(float(v) for v in orientation)

    def __eq__(self, other):
        return (isinstance(other, Pose) and self.position == other.position
                and self.orientation == other.orientation)

    def __repr__(self):
        return f"Pose(position={self.position}, orientation={self.orientation})"


def Pose_identity() -> Pose:
    return Pose()


class ContactPoint:
    """core/Link.h ContactPoint."""

    def __init__(self):
        self.depth = 0.0
        self.force = (0.0, 0.0, 0.0)
        self.torque = (0.0, 0.0, 0.0)
        self.normal = (0.0, 0.0, 0.0)
        self.position = (0.0, 0.0, 0.0)


class Contact:
    """core/Link.h Contact: the two scoped body names (``model::link``) and the contact points."""

    def __init__(self, body_a: str = "", body_b: str = "", points: Sequence[ContactPoint] = ()):
        self.body_a, self.body_b, self.points = body_a, body_b, list(points)


class Joint(abc.ABC):
    """scenario::core::Joint (pure virtual in the reference)."""


class Link(abc.ABC):
    """scenario::core::Link."""


class Model(abc.ABC):
    """scenario::core::Model."""


class World(abc.ABC):
    """scenario::core::World."""


def get_install_prefix() -> str:
    import os
    return os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# SWIG vector/array templates (core.i:33-51) are plain tuples here.
VectorD = tuple
VectorS = tuple
Array3d = tuple
Array4d = tuple
Array6d = tuple

__all__: List[str] = [n for n in dir() if not n.startswith("_")]
