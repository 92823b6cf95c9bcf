"""``scenario.gazebo``: ScenarI/O's Gazebo back-end API, served by the B200 engine.

Mirror of what bindings/gazebo/gazebo.i exports (GazeboSimulator, World, Model, Joint, Link, free functions),
with the semantics of cpp/scenario/gazebo/src/*.cpp. Every object is a *view* on one env of a batched
simulator (``env=0`` by default, which is the whole simulator when ``num_envs == 1``): getters copy a few
scalars from HBM, mutators write them; the physics runs in the CUDA kernels of ``b2sim``.

Error convention of the reference: mutators return ``bool`` (False plus a console message, e.g.
Joint.cpp:134-138), accessors raise ``RuntimeError`` (core.i:14-22).

Extensions over the reference (all optional keyword arguments): ``GazeboSimulator(..., num_envs, dtype,
device)`` and ``GazeboSimulator.get_world(name, env)``. Joint configuration (control mode, PID gains,
controller period) is shared by all envs of a simulator; state, targets and resets are per env.
"""
import math
import os
import random
import string
import sys
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence

import b2sim
from b2sim import _lib as _b2

from . import core
from .core import (Contact, ContactPoint, JointControlMode_force, JointControlMode_idle,  # noqa: F401
                   JointControlMode_invalid, JointControlMode_position, JointControlMode_position_interpolated,
                   JointControlMode_velocity, JointControlMode_velocity_follower_dart, JointLimit, JointType_ball,
                   JointType_fixed, JointType_invalid, JointType_prismatic, JointType_revolute, Limit, PID, Pose,
                   Pose_identity)

# enum class PhysicsEngine (gazebo/World.h:43-46) and Verbosity (gazebo/utils.h:40-47)
PhysicsEngine_dart = 0
Verbosity_suppress_all, Verbosity_error, Verbosity_warning, Verbosity_info, Verbosity_debug = range(5)

_verbosity = Verbosity_warning
_DBL_MAX = sys.float_info.max


def set_verbosity(level: int = Verbosity_warning) -> None:
    """gazebo/utils.h:75-83."""
    global _verbosity
    _verbosity = int(level)


def _log(level: int, tag: str, msg: str) -> None:
    if _verbosity >= level:
        print(f"[{tag}] {msg}", file=sys.stderr)


def _err(msg):
    _log(Verbosity_error, "Err", msg)


def _warn(msg):
    _log(Verbosity_warning, "Wrn", msg)


def _dbg(msg):
    _log(Verbosity_debug, "Dbg", msg)


# ---------------------------------------------------------------------------------------------------
# world registry: stands in for ECMSingleton (cpp/scenario/plugins/ECMProvider/ECMSingleton.cpp:66-201);
# python/gym_ignition/utils/scenario.py:47-57 asks it for the names of the live worlds.
# ---------------------------------------------------------------------------------------------------
class _WorldRegistry:
    def __init__(self):
        self._names: Dict[str, int] = {}

    def add(self, name: str):
        self._names[name] = self._names.get(name, 0) + 1

    def remove(self, name: str):
        if name in self._names:
            self._names[name] -= 1
            if self._names[name] <= 0:
                del self._names[name]

    def world_names(self) -> List[str]:
        return list(self._names)

    def valid(self, world_name: str = "") -> bool:
        return bool(self._names) if not world_name else world_name in self._names


_registry = _WorldRegistry()


def ECMSingleton_instance() -> _WorldRegistry:
    return _registry


_next_id = [1]


def _new_id() -> int:
    _next_id[0] += 1
    return _next_id[0]


def _names_or_all(names: Optional[Sequence[str]], default: List[str]) -> List[str]:
    return list(names) if names else list(default)


# ---------------------------------------------------------------------------------------------------
# Joint
# ---------------------------------------------------------------------------------------------------
class Joint(core.Joint):
    """scenario::gazebo::Joint (cpp/scenario/gazebo/src/Joint.cpp). Only 1-DoF joints exist in the engine,
    like in the reference (Joint.cpp:103-107)."""

    def __init__(self, model: "Model", index: int):
        self._model, self._j = model, index
        self._id = _new_id()
        self._history = None  # HistoryOfAppliedJointForces ring (helpers.h:84-115)

    # -- helpers --
    @property
    def _eng(self) -> b2sim.Simulator:
        return self._model._world._engine_checked()

    def _get(self, field) -> float:
        return self._eng.get_joint(self._model._mid, field, self._model._env, self._j)

    def _set(self, field, value) -> bool:
        try:
            self._eng.set_joint(self._model._mid, field, self._model._env, self._j, value)
            return True
        except b2sim.B2Error as e:
            _err(str(e))
            return False

    def _check_dof(self, dof):
        if dof != 0:
            raise RuntimeError(f"Joint '{self.name()}' does not have DoF#{dof}")  # DOFMismatch

    def _tables(self):
        return self._model._tables

    # -- identity --
    def id(self) -> int:
        return self._id

    def valid(self) -> bool:
        return self._model.valid()

    def to_gazebo(self) -> "Joint":
        return self

    def dofs(self) -> int:
        return 1

    def name(self, scoped: bool = False) -> str:
        n = self._model._info.joint_names[self._j]
        return f"{self._model.name()}::{n}" if scoped else n

    def type(self) -> int:
        return JointType_revolute if self._tables()["jtype"][self._j] == 2 else JointType_prismatic

    # -- control --
    def control_mode(self) -> int:
        return self._eng.control_mode(self._model._mid, self._j)

    def set_control_mode(self, mode: int) -> bool:
        try:
            self._eng.set_control_mode(self._model._mid, self._j, int(mode))
            return True
        except b2sim.B2Error as e:
            _err(str(e))
            return False

    def controller_period(self) -> float:
        return self._model.controller_period()

    def pid(self) -> PID:
        p = self._eng.pid(self._model._mid, self._j)
        out = PID(p.p, p.i, p.d)
        out.i_max, out.i_min = p.i_max, p.i_min
        out.cmd_max, out.cmd_min, out.cmd_offset = p.cmd_max, p.cmd_min, p.cmd_offset
        return out

    def set_pid(self, pid: PID) -> bool:
        try:
            self._eng.set_pid(self._model._mid, self._j, pid.p, pid.i, pid.d, pid.i_max, pid.i_min, pid.cmd_max,
                              pid.cmd_min, pid.cmd_offset)
            return True
        except b2sim.B2Error as e:
            _err(str(e))
            return False

    # -- force history (Joint.cpp:527-563) --
    def history_of_applied_joint_forces_enabled(self) -> bool:
        return self._history is not None

    def enable_history_of_applied_joint_forces(self, enable: bool = True, max_history_size: int = 100) -> bool:
        self._history = ([], int(max_history_size)) if enable else None
        return True

    def history_of_applied_joint_forces(self) -> tuple:
        if self._history is None:
            raise RuntimeError("The history of applied joint forces was not enabled")  # ComponentNotFound
        return tuple(self._history[0])

    def _push_history(self, value: float):
        if self._history is not None:
            ring, size = self._history
            ring.append(float(value))
            del ring[:-size]

    # -- constants --
    def coulomb_friction(self) -> float:
        return float(self._tables()["friction"][self._j])

    def viscous_friction(self) -> float:
        return float(self._tables()["damping"][self._j])

    def _set_friction(self, coulomb: float, viscous: float) -> bool:
        """Joint.cpp:259-311: allowed only while the parent model has just been created (helpers.cpp:131-157)."""
        if not self._model._parameters_editable():
            _err("The model has been already processed and its parameters cannot be modified")
            return False
        w = self._model._world
        rc = w._engine.lib.b2sim_set_joint_friction(w._engine.handle, self._model._mid, self._j, float(coulomb),
                                                    float(viscous))
        if rc < 0:
            _err(w._engine.lib.b2sim_last_error().decode())
            return False
        self._model._refresh_tables()
        return True

    def set_coulomb_friction(self, value: float) -> bool:
        return self._set_friction(value, -1.0)

    def set_viscous_friction(self, value: float) -> bool:
        return self._set_friction(-1.0, value)

    def position_limit(self, dof: int = 0) -> Limit:
        self._check_dof(dof)
        lo, hi = float(self._tables()["lower"][self._j]), float(self._tables()["upper"][self._j])
        return Limit(max(lo, -_DBL_MAX), min(hi, _DBL_MAX))

    def joint_position_limit(self) -> JointLimit:
        lim = self.position_limit()
        return JointLimit([lim.min], [lim.max])

    def max_generalized_force(self, dof: int = 0) -> float:
        self._check_dof(dof)
        f = self._model._effort[self._j]
        return min(f, _DBL_MAX)

    def set_max_generalized_force(self, max_force: float, dof: int = 0) -> bool:
        if dof != 0:
            _err(f"Joint '{self.name()}' does not have DoF#{dof}")
            return False
        if not self._model._parameters_editable():  # helpers.cpp:131-157, Joint.cpp:908-940
            _err("The model has been already processed and its parameters cannot be modified")
            return False
        w = self._model._world
        rc = w._engine.lib.b2sim_set_max_generalized_force(w._engine.handle, self._model._mid, self._j, float(max_force))
        if rc < 0:
            _err(w._engine.lib.b2sim_last_error().decode())
            return False
        self._model._effort[self._j] = float(max_force)
        return True

    def joint_max_generalized_force(self) -> tuple:
        return (self.max_generalized_force(),)

    def set_joint_max_generalized_force(self, max_force: Sequence[float]) -> bool:
        if len(max_force) != 1:
            _err(f"Wrong number of elements (joint_dofs={self.dofs()})")
            return False
        return self.set_max_generalized_force(max_force[0])

    # -- state --
    def position(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_POSITION)

    def velocity(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_VELOCITY)

    def acceleration(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_ACCELERATION)

    def generalized_force(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_FORCE)

    def joint_position(self) -> tuple:
        return (self.position(),)

    def joint_velocity(self) -> tuple:
        return (self.velocity(),)

    def joint_acceleration(self) -> tuple:
        return (self.acceleration(),)

    def joint_generalized_force(self) -> tuple:
        return (self.generalized_force(),)

    # -- targets --
    def set_position_target(self, position: float, dof: int = 0) -> bool:
        return dof == 0 and self._set(_b2.FIELD_POSITION_TARGET, position)

    def set_velocity_target(self, velocity: float, dof: int = 0) -> bool:
        return dof == 0 and self._set(_b2.FIELD_VELOCITY_TARGET, velocity)

    def set_acceleration_target(self, acceleration: float, dof: int = 0) -> bool:
        # Joint.cpp:731-772: accepted in PositionInterpolated / Idle / Force; consumed by custom controllers only
        return dof == 0 and self._set(_b2.FIELD_ACCELERATION_TARGET, acceleration)

    def set_generalized_force_target(self, force: float, dof: int = 0) -> bool:
        if dof != 0:
            _err(f"Joint '{self.name()}' does not have DoF#{dof}")
            return False
        if abs(force) > self.max_generalized_force():
            _warn("The force target is higher than the limit. The physics engine might clip it.")
        return self._set(_b2.FIELD_FORCE_TARGET, force)

    def position_target(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_POSITION_TARGET)

    def velocity_target(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_VELOCITY_TARGET)

    def acceleration_target(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_ACCELERATION_TARGET)

    def generalized_force_target(self, dof: int = 0) -> float:
        self._check_dof(dof)
        return self._get(_b2.FIELD_FORCE_TARGET)

    def set_joint_position_target(self, position: Sequence[float]) -> bool:
        return len(position) == 1 and self.set_position_target(position[0])

    def set_joint_velocity_target(self, velocity: Sequence[float]) -> bool:
        return len(velocity) == 1 and self.set_velocity_target(velocity[0])

    def set_joint_acceleration_target(self, acceleration: Sequence[float]) -> bool:
        return len(acceleration) == 1 and self.set_acceleration_target(acceleration[0])

    def set_joint_generalized_force_target(self, force: Sequence[float]) -> bool:
        return len(force) == 1 and self.set_generalized_force_target(force[0])

    def joint_position_target(self) -> tuple:
        return (self.position_target(),)

    def joint_velocity_target(self) -> tuple:
        return (self.velocity_target(),)

    def joint_acceleration_target(self) -> tuple:
        return (self.acceleration_target(),)

    def joint_generalized_force_target(self) -> tuple:
        return (self.generalized_force_target(),)

    # -- resets: consumed by the next run, paused or not (Physics.cpp:1330-1375) --
    def reset_position(self, position: float = 0.0, dof: int = 0) -> bool:
        return dof == 0 and self._set(_b2.FIELD_POSITION_RESET, position)

    def reset_velocity(self, velocity: float = 0.0, dof: int = 0) -> bool:
        return dof == 0 and self._set(_b2.FIELD_VELOCITY_RESET, velocity)

    def reset(self, position: float = 0.0, velocity: float = 0.0, dof: int = 0) -> bool:
        return self.reset_position(position, dof) and self.reset_velocity(velocity, dof)

    def reset_joint_position(self, position: Sequence[float]) -> bool:
        return len(position) == 1 and self.reset_position(position[0])

    def reset_joint_velocity(self, velocity: Sequence[float]) -> bool:
        return len(velocity) == 1 and self.reset_velocity(velocity[0])

    def reset_joint(self, position: Sequence[float], velocity: Sequence[float]) -> bool:
        return self.reset_joint_position(position) and self.reset_joint_velocity(velocity)


# ---------------------------------------------------------------------------------------------------
# Link
# ---------------------------------------------------------------------------------------------------
def _quat_to_R(q):
    w, x, y, z = q
    return [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]


class Link(core.Link):
    """scenario::gazebo::Link (cpp/scenario/gazebo/src/Link.cpp)."""

    def __init__(self, model: "Model", index: int):
        self._model, self._l = model, index
        self._id = _new_id()
        self._contacts_enabled = False

    def id(self) -> int:
        return self._id

    def valid(self) -> bool:
        return self._model.valid()

    def to_gazebo(self) -> "Link":
        return self

    def name(self, scoped: bool = False) -> str:
        n = self._model._info.link_names[self._l]
        return f"{self._model.name()}::{n}" if scoped else n

    def mass(self) -> float:
        return float(self._model._tables["link_mass"][self._l])

    def _pose(self):
        m = self._model
        return m._world._engine_checked().link_pose(m._mid, m._env, self._l)

    def position(self) -> tuple:
        return tuple(self._pose()[:3])

    def orientation(self) -> tuple:
        return tuple(self._pose()[3:])  # wxyz, helpers.cpp:159-174

    def _world_twist(self):
        """6-vector [linear; angular] of the link frame origin, world orientation: J(q) dq."""
        import torch
        m = self._model
        eng = m._world._engine_checked()
        nq = m.dofs()
        if m._info.kind == _b2.KIND_FREE:
            st = eng.base_state(m._mid, m._env)
            w = st[10:13]
            p_link, p_base = self.position(), st[0:3]
            r = [p_link[k] - p_base[k] for k in range(3)]
            v = [st[7] + w[1] * r[2] - w[2] * r[1], st[8] + w[2] * r[0] - w[0] * r[2], st[9] + w[0] * r[1] - w[1] * r[0]]
            return v + list(w)
        if nq == 0:
            return [0.0] * 6
        tdt = torch.float64 if eng.dtype == "float64" else torch.float32
        J = torch.empty((eng.num_envs, 6 * nq), dtype=tdt, device=torch.device("cuda", eng.device))
        eng.kindyn(m._mid, self._l, None, None, J)
        dq = eng.tensor(m._mid, _b2.BUF_STATE)[m._env, nq:].double()
        return (J[m._env].double().view(6, nq) @ dq).tolist()

    def world_linear_velocity(self) -> tuple:
        return tuple(self._world_twist()[:3])

    def world_angular_velocity(self) -> tuple:
        return tuple(self._world_twist()[3:])

    def _to_body(self, v):
        R = _quat_to_R(self.orientation())  # body = R^T world, Physics.cpp:2020-2079
        return tuple(sum(R[k][i] * v[k] for k in range(3)) for i in range(3))

    def body_linear_velocity(self) -> tuple:
        return self._to_body(self.world_linear_velocity())

    def body_angular_velocity(self) -> tuple:
        return self._to_body(self.world_angular_velocity())

    def _world_accel(self):
        """6-vector [linear; angular]: classical acceleration of the link origin, world orientation."""
        import torch
        m = self._model
        eng = m._world._engine_checked()
        if m.dofs() == 0:
            if m._info.kind == _b2.KIND_FREE:
                # rigid body: a_link = a_base + alpha x r + w x (w x r); the base acceleration is the velocity change of
                # the last step / dt, constraint impulses included (B2_BUF_BASE_ACCEL)
                acc = eng.tensor(m._mid, _b2.BUF_BASE_ACCEL)[m._env].double().tolist()
                st = eng.base_state(m._mid, m._env)
                w, al = st[10:13], acc[3:6]
                p_link = self.position()
                r = [p_link[k] - st[k] for k in range(3)]
                cr = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
                axr, wwr = cr(al, r), cr(w, cr(w, r))
                return [acc[k] + axr[k] + wwr[k] for k in range(3)] + list(al)
            return [0.0] * 6
        tdt = torch.float64 if eng.dtype == "float64" else torch.float32
        out = torch.empty((eng.num_envs, 6), dtype=tdt, device=torch.device("cuda", eng.device))
        eng.link_motion(m._mid, self._l, None, out)
        return out[m._env].tolist()

    def world_linear_acceleration(self) -> tuple:
        return tuple(self._world_accel()[:3])

    def world_angular_acceleration(self) -> tuple:
        return tuple(self._world_accel()[3:])

    def body_linear_acceleration(self) -> tuple:
        return self._to_body(self.world_linear_acceleration())

    def body_angular_acceleration(self) -> tuple:
        return self._to_body(self.world_angular_acceleration())

    # contacts (Link.cpp:296-482). The engine simulates contacts between free-floating bodies and static shapes;
    # they are *reported* only for links with contact detection enabled, like in the reference (Appendix A.13).
    def contacts_enabled(self) -> bool:
        return self._contacts_enabled

    def enable_contact_detection(self, enable: bool) -> bool:
        self._contacts_enabled = bool(enable)
        return True

    def in_contact(self) -> bool:
        return len(self.contacts()) > 0

    def contacts(self) -> tuple:
        """One Contact per touching body pair, seen from this link: body_a is this link; normals and forces of
        pairs where this link is the second body are flipped (Physics.cpp:2498-2529, helpers.cpp:191-275)."""
        if not self._contacts_enabled:
            return ()
        m = self._model
        eng = m._world._engine_checked()
        merged = {}
        for (ma, la, mb, lb, pos, normal, depth, force) in eng.contacts(m._env):
            if (ma, la) == (m._mid, self._l):
                other, sign = (mb, lb), 1.0
            elif (mb, lb) == (m._mid, self._l):
                other, sign = (ma, la), -1.0
            else:
                continue
            point = ContactPoint()
            point.depth = depth
            point.position = pos
            point.normal = tuple(sign * v for v in normal)
            point.force = tuple(sign * v for v in force)
            point.torque = (0.0, 0.0, 0.0)  # Physics.cpp:2531-2532
            merged.setdefault(other, []).append(point)
        out = []
        for (om, ol), points in merged.items():
            other_name = f"{eng.lib.b2sim_model_name(eng.handle, om).decode()}::{eng.info(om).link_names[ol]}"
            out.append(Contact(self.name(scoped=True), other_name, points))
        return tuple(out)

    def contact_wrench(self) -> tuple:
        """Sum of the contact forces and of their moments about the link origin, world orientation (Link.cpp:436-482)."""
        origin = self.position()
        f = [0.0, 0.0, 0.0]
        t = [0.0, 0.0, 0.0]
        for contact in self.contacts():
            for p in contact.points:
                r = [p.position[k] - origin[k] for k in range(3)]
                for k in range(3):
                    f[k] += p.force[k]
                t[0] += r[1] * p.force[2] - r[2] * p.force[1]
                t[1] += r[2] * p.force[0] - r[0] * p.force[2]
                t[2] += r[0] * p.force[1] - r[1] * p.force[0]
        return tuple(f + t)

    # external wrenches with duration (Link.cpp:484-557): world frame, applied at the link origin
    def apply_world_wrench(self, force, torque, duration: float = 0.0) -> bool:
        m = self._model
        try:
            m._world._engine_checked().apply_link_wrench(m._mid, m._env, self._l, list(force) + list(torque), duration)
            return True
        except b2sim.B2Error as e:
            _err(str(e))
            return False

    def apply_world_wrench_to_com(self, force, torque, duration: float = 0.0) -> bool:
        """Link.cpp:529-557: the wrench commands act at the link origin, so a force through the centre of mass adds
        the torque (W_R_L L_o_com) x force."""
        m = self._model
        com = m._link_com(self._l)
        R = _quat_to_R(self.orientation())
        r = [sum(R[i][k] * com[k] for k in range(3)) for i in range(3)]
        f, t = list(force), list(torque)
        t[0] += r[1] * f[2] - r[2] * f[1]
        t[1] += r[2] * f[0] - r[0] * f[2]
        t[2] += r[0] * f[1] - r[1] * f[0]
        return self.apply_world_wrench(f, t, duration)

    apply_world_wrench_to_co_m = apply_world_wrench_to_com  # the name SWIG's %(undercase)s gives applyWorldWrenchToCoM

    def apply_world_force(self, force, duration: float = 0.0) -> bool:
        return self.apply_world_wrench(force, (0.0, 0.0, 0.0), duration)

    def apply_world_torque(self, torque, duration: float = 0.0) -> bool:
        return self.apply_world_wrench((0.0, 0.0, 0.0), torque, duration)


# ---------------------------------------------------------------------------------------------------
# Model
# ---------------------------------------------------------------------------------------------------
class Model(core.Model):
    """scenario::gazebo::Model (cpp/scenario/gazebo/src/Model.cpp)."""

    def __init__(self, world: "World", mid: int, name: str, pose: Pose):
        self._world, self._mid, self._name, self._pose0 = world, mid, name, pose
        self._env = world._env
        self._info = world._engine.info(mid)
        self._tables = self._info.tables()
        self._effort = [float(v) for v in self._tables["effort"]]
        self._joints: Dict[str, Joint] = {}
        self._links: Dict[str, Link] = {}
        self._acc_targets: Dict[int, float] = {}
        self._base_targets: Dict[str, tuple] = {}
        self._row_cache: Dict[int, tuple] = {}
        self._id = _new_id()
        self._removed = False
        self._timestamp_ns = world._time_ns  # components::Timestamp, Model.cpp:143-150
        self._self_collisions = False

    def _refresh_tables(self):
        self._tables = self._info.tables()
        return self._tables

    def _link_com(self, link: int):
        return [float(v) for v in self._tables["link_com"][link]]

    # -- identity --
    def id(self) -> int:
        return self._id

    def valid(self) -> bool:
        return not self._removed and self._world.valid()

    def to_gazebo(self) -> "Model":
        return self

    def name(self) -> str:
        return self._name

    def dofs(self) -> int:
        return self._info.dofs

    def nr_of_links(self) -> int:
        return len(self._info.link_names)

    def nr_of_joints(self) -> int:
        return len(self._info.joint_names)

    def total_mass(self) -> float:
        return float(self._tables["total_mass"])

    def link_names(self, scoped: bool = False) -> tuple:
        return tuple(f"{self._name}::{n}" if scoped else n for n in self._info.link_names)

    def joint_names(self, scoped: bool = False) -> tuple:
        return tuple(f"{self._name}::{n}" if scoped else n for n in self._info.joint_names)

    def _parameters_editable(self) -> bool:
        return self._world._time_ns == self._timestamp_ns

    def get_link(self, link_name: str) -> Link:
        if link_name not in self._links:
            if link_name not in self._info.link_names:
                raise RuntimeError(f"Link '{link_name}' not found")  # LinkNotFound, Model.cpp:436-438
            self._links[link_name] = Link(self, self._info.link_names.index(link_name))
        return self._links[link_name]

    def get_joint(self, joint_name: str) -> Joint:
        if joint_name not in self._joints:
            if joint_name not in self._info.joint_names:
                raise RuntimeError(f"Joint '{joint_name}' not found")  # JointNotFound, Model.cpp:462-464
            self._joints[joint_name] = Joint(self, self._info.joint_names.index(joint_name))
        return self._joints[joint_name]

    def links(self, link_names: Sequence[str] = ()) -> tuple:
        return tuple(self.get_link(n) for n in _names_or_all(link_names, self._info.link_names))

    def joints(self, joint_names: Sequence[str] = ()) -> tuple:
        return tuple(self.get_joint(n) for n in _names_or_all(joint_names, self._info.joint_names))

    # -- controller --
    def controller_period(self) -> float:
        return self._world._engine_checked().controller_period(self._mid)

    def set_controller_period(self, period: float) -> bool:
        try:
            self._world._engine_checked().set_controller_period(self._mid, float(period))
            return True
        except b2sim.B2Error as e:
            _err(str(e))
            return False

    def set_joint_control_mode(self, mode: int, joint_names: Sequence[str] = ()) -> bool:
        return all([j.set_control_mode(mode) for j in self.joints(joint_names)])

    # -- history of applied forces (Model.cpp:604-672) --
    def enable_history_of_applied_joint_forces(self, enable: bool = True, max_history_size_per_joint: int = 100,
                                               joint_names: Sequence[str] = ()) -> bool:
        return all([j.enable_history_of_applied_joint_forces(enable, max_history_size_per_joint)
                    for j in self.joints(joint_names)])

    def history_of_applied_joint_forces_enabled(self, joint_names: Sequence[str] = ()) -> bool:
        return all(j.history_of_applied_joint_forces_enabled() for j in self.joints(joint_names))

    def history_of_applied_joint_forces(self, joint_names: Sequence[str] = ()) -> tuple:
        joints = self.joints(joint_names)
        rings = [j.history_of_applied_joint_forces() for j in joints]
        n = min(len(r) for r in rings) if rings else 0
        out = []
        for k in range(n):  # time-major: all joints at step k, then step k+1 (Model.cpp:653-670)
            out.extend(r[len(r) - n + k] for r in rings)
        return tuple(out)

    # -- contacts (not built yet) --
    def contacts_enabled(self) -> bool:
        return all(l.contacts_enabled() for l in self.links())

    def enable_contacts(self, enable: bool = True) -> bool:
        return all([l.enable_contact_detection(enable) for l in self.links()])

    def self_collisions_enabled(self) -> bool:
        return self._self_collisions

    def enable_self_collisions(self, enable: bool = True) -> bool:
        if not self._parameters_editable():
            _err("The model has been already processed and its parameters cannot be modified")
            return False
        self._self_collisions = bool(enable)
        return True

    def links_in_contact(self) -> tuple:
        return tuple(l.name() for l in self.links() if l.in_contact())

    def contacts(self, link_names: Sequence[str] = ()) -> tuple:
        out = []
        for link in self.links(link_names):
            out.extend(link.contacts())
        return tuple(out)

    # -- vectorised joint access, serialised in the caller's joint order (Model.cpp:756-794,1249-1267) --
    def _joint_indices(self, joint_names) -> List[int]:
        names = _names_or_all(joint_names, self._info.joint_names)
        idx = []
        for n in names:
            if n not in self._info.joint_names:
                raise RuntimeError(f"Joint '{n}' not found")
            idx.append(self._info.joint_names.index(n))
        return idx

    def _row(self, which) -> List[float]:
        """One env's row of a state buffer. The joint state only changes inside run() (resets are deferred to it,
        Physics.cpp:1330-1375), so the row is fetched once per run: a Task reads it several times per env.step
        (get_observation in step and again in is_done), each fetch being a device -> host round trip."""
        eng = self._world._engine_checked()
        if self.dofs() == 0:
            return []
        epoch = getattr(eng, "run_count", 0)
        hit = self._row_cache.get(which)
        if hit is not None and hit[0] == epoch:
            return hit[1]
        row = eng.tensor(self._mid, which)[self._env].tolist()
        self._row_cache[which] = (epoch, row)
        return row

    def joint_positions(self, joint_names: Sequence[str] = ()) -> tuple:
        row = self._row(_b2.BUF_STATE)
        return tuple(row[j] for j in self._joint_indices(joint_names))

    def joint_velocities(self, joint_names: Sequence[str] = ()) -> tuple:
        row, nq = self._row(_b2.BUF_STATE), self.dofs()
        return tuple(row[nq + j] for j in self._joint_indices(joint_names))

    def joint_accelerations(self, joint_names: Sequence[str] = ()) -> tuple:
        row = self._row(_b2.BUF_ACCELERATION)
        return tuple(row[j] for j in self._joint_indices(joint_names))

    def joint_generalized_forces(self, joint_names: Sequence[str] = ()) -> tuple:
        return tuple(j.generalized_force() for j in self.joints(joint_names))

    def joint_limits(self, joint_names: Sequence[str] = ()) -> JointLimit:
        lims = [j.position_limit() for j in self.joints(joint_names)]
        return JointLimit([l.min for l in lims], [l.max for l in lims])

    def _set_many(self, setter: str, values: Sequence[float], joint_names: Sequence[str]) -> bool:
        joints = self.joints(joint_names)
        if len(values) != len(joints):
            _err("The size of the values does not match the considered joint's DOFs")  # Model.cpp:1281-1286
            return False
        return all([getattr(j, setter)(float(v)) for j, v in zip(joints, values)])

    def set_joint_position_targets(self, positions, joint_names: Sequence[str] = ()) -> bool:
        return self._set_many("set_position_target", positions, joint_names)

    def set_joint_velocity_targets(self, velocities, joint_names: Sequence[str] = ()) -> bool:
        return self._set_many("set_velocity_target", velocities, joint_names)

    def set_joint_acceleration_targets(self, accelerations, joint_names: Sequence[str] = ()) -> bool:
        return self._set_many("set_acceleration_target", accelerations, joint_names)

    def set_joint_generalized_force_targets(self, forces, joint_names: Sequence[str] = ()) -> bool:
        return self._set_many("set_generalized_force_target", forces, joint_names)

    def joint_position_targets(self, joint_names: Sequence[str] = ()) -> tuple:
        return tuple(j.position_target() for j in self.joints(joint_names))

    def joint_velocity_targets(self, joint_names: Sequence[str] = ()) -> tuple:
        return tuple(j.velocity_target() for j in self.joints(joint_names))

    def joint_acceleration_targets(self, joint_names: Sequence[str] = ()) -> tuple:
        return tuple(j.acceleration_target() for j in self.joints(joint_names))

    def joint_generalized_force_targets(self, joint_names: Sequence[str] = ()) -> tuple:
        return tuple(j.generalized_force_target() for j in self.joints(joint_names))

    def reset_joint_positions(self, positions, joint_names: Sequence[str] = ()) -> bool:
        return self._set_many("reset_position", positions, joint_names)  # Model.cpp:230-241

    def reset_joint_velocities(self, velocities, joint_names: Sequence[str] = ()) -> bool:
        return self._set_many("reset_velocity", velocities, joint_names)  # Model.cpp:243-254

    # -- base (fixed-base models: the base frame is the first link, at the insertion pose) --
    def base_frame(self) -> str:
        return self._info.link_names[0] if self._info.link_names else ""

    def base_position(self) -> tuple:
        return self.get_link(self.base_frame()).position() if self._info.link_names else self._pose0.position

    def base_orientation(self) -> tuple:
        return self.get_link(self.base_frame()).orientation() if self._info.link_names else self._pose0.orientation

    def _free(self) -> bool:
        return self._info.kind == _b2.KIND_FREE

    def _base_state(self):
        return self._world._engine_checked().base_state(self._mid, self._env)

    def base_world_linear_velocity(self) -> tuple:
        return tuple(self._base_state()[7:10]) if self._free() else (0.0, 0.0, 0.0)

    def base_world_angular_velocity(self) -> tuple:
        return tuple(self._base_state()[10:13]) if self._free() else (0.0, 0.0, 0.0)

    def _world_to_body(self, v):
        R = _quat_to_R(self.base_orientation())
        return tuple(sum(R[k][i] * v[k] for k in range(3)) for i in range(3))

    def base_body_linear_velocity(self) -> tuple:
        return self._world_to_body(self.base_world_linear_velocity())

    def base_body_angular_velocity(self) -> tuple:
        return self._world_to_body(self.base_world_angular_velocity())

    # base resets are consumed by the next run (WorldPoseCmd / WorldVelocityCmd, Model.cpp:256-377)
    def _set_base(self, values, velocity: bool) -> bool:
        if not self._free():
            _err("the model has a fixed base: its pose cannot be reset")
            return False
        try:
            self._world._engine_checked().set_base(self._mid, self._env, values, velocity)
            return True
        except b2sim.B2Error as e:
            _err(str(e))
            return False

    def reset_base_pose(self, position=(0.0, 0.0, 0.0), orientation=(1.0, 0.0, 0.0, 0.0)) -> bool:
        return self._set_base(list(position) + list(orientation), False)

    def reset_base_position(self, position=(0.0, 0.0, 0.0)) -> bool:
        return self.reset_base_pose(position, self.base_orientation())

    def reset_base_orientation(self, orientation=(1.0, 0.0, 0.0, 0.0)) -> bool:
        return self.reset_base_pose(self.base_position(), orientation)

    def reset_base_world_velocity(self, linear=(0.0, 0.0, 0.0), angular=(0.0, 0.0, 0.0)) -> bool:
        return self._set_base(list(linear) + list(angular), True)

    def reset_base_world_linear_velocity(self, linear=(0.0, 0.0, 0.0)) -> bool:
        return self.reset_base_world_velocity(linear, self.base_world_angular_velocity())

    def reset_base_world_angular_velocity(self, angular=(0.0, 0.0, 0.0)) -> bool:
        return self.reset_base_world_velocity(self.base_world_linear_velocity(), angular)

    # -- base targets (Model.cpp:1077-1247): plain components, read back by custom controllers only
    # (ControllerRunner.cpp:319-369); physics never consumes them. Getters raise when the component was never set,
    # like utils::getExistingComponentData.
    def _base_target(self, key):
        try:
            return self._base_targets[key]
        except KeyError:
            raise RuntimeError(f"Component '{key}' not found in the model '{self._name}'")

    def set_base_pose_target(self, position, orientation) -> bool:
        self._base_targets["BasePoseTarget"] = (tuple(float(v) for v in position), tuple(float(v) for v in orientation))
        return True

    def set_base_position_target(self, position) -> bool:
        # the reference starts from Pose3d::Zero when the component is missing (Model.cpp:1091-1093)
        _, quat = self._base_targets.get("BasePoseTarget", ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0, 0.0)))
        return self.set_base_pose_target(position, quat)

    def set_base_orientation_target(self, orientation) -> bool:
        pos, _ = self._base_targets.get("BasePoseTarget", ((0.0, 0.0, 0.0), (1.0, 0.0, 0.0, 0.0)))
        return self.set_base_pose_target(pos, orientation)

    def set_base_world_velocity_target(self, linear, angular) -> bool:
        return self.set_base_world_linear_velocity_target(linear) and self.set_base_world_angular_velocity_target(angular)

    def set_base_world_linear_velocity_target(self, linear) -> bool:
        self._base_targets["BaseWorldLinearVelocityTarget"] = tuple(float(v) for v in linear)
        return True

    def set_base_world_angular_velocity_target(self, angular) -> bool:
        self._base_targets["BaseWorldAngularVelocityTarget"] = tuple(float(v) for v in angular)
        return True

    def set_base_world_linear_acceleration_target(self, linear) -> bool:
        self._base_targets["BaseWorldLinearAccelerationTarget"] = tuple(float(v) for v in linear)
        return True

    def set_base_world_angular_acceleration_target(self, angular) -> bool:
        self._base_targets["BaseWorldAngularAccelerationTarget"] = tuple(float(v) for v in angular)
        return True

    def base_position_target(self) -> tuple:
        return self._base_target("BasePoseTarget")[0]

    def base_orientation_target(self) -> tuple:
        return self._base_target("BasePoseTarget")[1]

    def base_world_linear_velocity_target(self) -> tuple:
        return self._base_target("BaseWorldLinearVelocityTarget")

    def base_world_angular_velocity_target(self) -> tuple:
        return self._base_target("BaseWorldAngularVelocityTarget")

    def base_world_linear_acceleration_target(self) -> tuple:
        return self._base_target("BaseWorldLinearAccelerationTarget")

    def base_world_angular_acceleration_target(self) -> tuple:
        return self._base_target("BaseWorldAngularAccelerationTarget")

    def insert_model_plugin(self, lib_name: str, class_name: str, context: str = "") -> bool:
        """Model.cpp:190-228. JointController is built into the step kernel; ControllerRunner accepts the
        ComputedTorqueFixedBase controller (ControllersFactory.cpp:74-126 parses the same <controller> context)."""
        if class_name.endswith("JointController"):
            return True
        if class_name.endswith("ControllerRunner"):
            try:
                root = ET.fromstring(context)
                ctrl = root if root.tag == "controller" else root.find("controller")
                if ctrl is None or ctrl.get("name") != "ComputedTorqueFixedBase":
                    _err("Only the ComputedTorqueFixedBase controller is available")
                    return False
                floats = lambda tag: [float(v) for v in ctrl.find(tag).text.split()]
                kp, kd = floats("kp"), floats("kd")
                joints = ctrl.find("joints").text.split()
                gravity = floats("gravity") if ctrl.find("gravity") is not None else [0.0, 0.0, -9.80665]
            except Exception as e:  # noqa: BLE001
                _err(f"Failed to parse the controller context: {e}")
                return False
            if set(joints) != set(self._info.joint_names) or len(kp) != len(joints) or len(kd) != len(joints):
                _err("Controlling only a subset of joints is not yet supported")  # ComputedTorqueFixedBase.cpp:147-151
                return False
            order = [joints.index(n) for n in self._info.joint_names]
            try:
                self._world._engine_checked().set_computed_torque(self._mid, [kp[i] for i in order],
                                                                  [kd[i] for i in order], gravity)
                return True
            except b2sim.B2Error as e:
                _err(str(e))
                return False
        _err(f"model plugin '{class_name}' is not available in the B200 engine")
        return False

    # called by World.run: JointForceCmd history is appended on unpaused steps only (Physics.cpp:2085-2112)
    def _record_history(self, pre_run_cmds: List[float], iterations: int):
        if not any(j._history is not None for j in self._joints.values()):
            return
        eng = self._world._engine_checked()
        nq = self.dofs()
        pid_state = eng.tensor(self._mid, _b2.BUF_PID_STATE)[self._env].tolist()
        for name, j in self._joints.items():
            if j._history is None:
                continue
            mode = eng.control_mode(self._mid, j._j)
            for it in range(iterations):
                if mode in (JointControlMode_position, JointControlMode_velocity):
                    j._push_history(pid_state[3 * j._j + 2])
                else:
                    j._push_history(pre_run_cmds[j._j] if it == 0 else 0.0)


# ---------------------------------------------------------------------------------------------------
# World
# ---------------------------------------------------------------------------------------------------
class World(core.World):
    """scenario::gazebo::World (cpp/scenario/gazebo/src/World.cpp)."""

    def __init__(self, simulator: "GazeboSimulator", name: str, env: int = 0, shared: "World" = None):
        self._sim, self._name, self._env = simulator, name, env
        self._shared = shared if shared is not None else self  # env views share the engine and bookkeeping
        self._id = _new_id()
        if shared is None:
            self._engine: Optional[b2sim.Simulator] = None
            self._models: Dict[str, Model] = {}
            self._pending_removal: List[str] = []
            self._physics_loaded = False
            self._time_ns = 0
            self._gravity = (0.0, 0.0, -9.8)
        self._views: Dict[str, Model] = {}

    def __getattr__(self, item):
        # env views delegate the shared bookkeeping to the env-0 world
        if item in ("_engine", "_models", "_pending_removal", "_physics_loaded", "_time_ns", "_gravity"):
            shared = object.__getattribute__(self, "_shared")
            if shared is not self:
                return getattr(shared, item)
        raise AttributeError(item)

    def _engine_checked(self) -> b2sim.Simulator:
        if self._engine is None:
            raise RuntimeError("The simulator was not initialized or was closed")
        return self._engine

    def id(self) -> int:
        return self._id

    def valid(self) -> bool:
        return self._shared._engine is not None

    def to_gazebo(self) -> "World":
        return self

    def name(self) -> str:
        return self._name

    def time(self) -> float:
        return self._shared._time_ns / 1e9  # components::SimulatedTime, written by Physics (Physics.cpp:656-666)

    def gravity(self) -> tuple:
        return tuple(self._shared._gravity)

    def set_gravity(self, gravity: Sequence[float]) -> bool:
        sh = self._shared
        if sh._physics_loaded and sh._time_ns != 0:  # World.cpp:301-319
            _err("Physics has already advanced: the gravity cannot be changed any more")
            return False
        try:
            sh._engine_checked().set_gravity([float(g) for g in gravity])
        except b2sim.B2Error as e:
            _err(str(e))
            return False
        sh._gravity = tuple(float(g) for g in gravity)
        return True

    def set_physics_engine(self, engine: int = PhysicsEngine_dart) -> bool:
        if engine != PhysicsEngine_dart:
            _err("Unsupported physics engine")
            return False
        self._shared._physics_loaded = True  # World.cpp:273-291: loads the Physics system
        return True

    def insert_world_plugin(self, lib_name: str, class_name: str, context: str = "") -> bool:
        if class_name.endswith("Physics"):
            return self.set_physics_engine(PhysicsEngine_dart)
        _err(f"world plugin '{class_name}' is not available in the B200 engine")
        return False

    def model_names(self) -> tuple:
        return tuple(self._shared._models)

    def get_model(self, model_name: str) -> Model:
        sh = self._shared
        if model_name not in sh._models:
            raise RuntimeError(f"Model '{model_name}' not found")  # ModelNotFound, World.cpp:382-384
        base = sh._models[model_name]
        if self is sh:
            return base
        if model_name not in self._views or self._views[model_name]._mid != base._mid:
            view = Model.__new__(Model)
            view.__dict__.update(base.__dict__)
            view._world, view._env = self, self._env
            view._joints, view._links, view._acc_targets = {}, {}, {}
            self._views[model_name] = view
        return self._views[model_name]

    # -- insertion / removal (World.cpp:394-453) --
    def insert_model(self, model_file: str, pose: Pose = None, override_model_name: str = "") -> bool:
        return self.insert_model_from_file(model_file, pose, override_model_name)

    def insert_model_from_file(self, path: str, pose: Pose = None, override_model_name: str = "") -> bool:
        if not os.path.isfile(path):
            _err(f"Failed to find model file '{path}'")
            return False
        with open(path, "r") as f:
            return self.insert_model_from_string(f.read(), pose, override_model_name)

    def insert_model_from_string(self, sdf_string: str, pose: Pose = None, override_model_name: str = "") -> bool:
        sh = self._shared
        pose = pose if pose is not None else Pose_identity()
        try:
            eng = sh._engine_checked()
        except RuntimeError as e:
            _err(str(e))
            return False
        try:
            info = b2sim.ModelInfo.from_string(sdf_string)
        except b2sim.B2Error as e:
            _err(str(e))
            return False
        name = override_model_name or info.name
        if name in sh._models:  # World.cpp:86-93
            _err(f"Failed to insert model '{name}': another entity with the same name already exists")
            return False
        try:
            mid = eng.insert_model(sdf_string, list(pose.position) + list(pose.orientation), name)
        except b2sim.B2Error as e:
            _err(str(e))
            return False
        sh._models[name] = Model(sh, mid, name, pose)
        return True

    def remove_model(self, model_name: str) -> bool:
        sh = self._shared
        if model_name not in sh._models:
            _err(f"Model '{model_name}' not found in the world")
            return False
        if model_name not in sh._pending_removal:
            sh._pending_removal.append(model_name)  # processed by the next run (World.cpp:431-453)
        return True

    def _process_removals(self):
        for name in self._pending_removal:
            model = self._models.pop(name, None)
            if model is not None:
                model._removed = True
                try:
                    self._engine.remove_model(model._mid)
                except b2sim.B2Error as e:
                    _err(str(e))
        self._pending_removal.clear()

    def _run(self, paused: bool, server_time_ns: int):
        """One GazeboSimulator::run for this world (env-0 object only)."""
        if self._engine is None:
            return
        if self._physics_loaded:
            models = [m for m in self._models.values() if m.dofs() > 0]
            pre = {}
            if not paused:
                for m in models:
                    views = [m] + [w._views[m._name] for w in self._sim._env_views(self) if m._name in w._views]
                    for v in views:
                        if any(j._history is not None for j in v._joints.values()):
                            pre[id(v)] = (v, self._engine.tensor(v._mid, _b2.BUF_FORCE_CMD)[v._env].tolist())
            self._engine.run(paused)
            for v, cmds in pre.values():
                v._record_history(cmds, self._sim.steps_per_run())
            self._time_ns = server_time_ns  # "physics catches up", tests/test_scenario/test_world.py:177-188
        self._process_removals()


# ---------------------------------------------------------------------------------------------------
# GazeboSimulator
# ---------------------------------------------------------------------------------------------------
class GazeboSimulator:
    """scenario::gazebo::GazeboSimulator (cpp/scenario/gazebo/src/GazeboSimulator.cpp)."""

    def __init__(self, step_size: float = 0.001, rtf: float = 1.0, steps_per_run: int = 1, num_envs: int = 1,
                 dtype: str = "float64", device: int = 0):
        self._step_size, self._rtf, self._steps = float(step_size), float(rtf), int(steps_per_run)
        self._num_envs, self._dtype, self._device = int(num_envs), dtype, int(device)
        self._worlds: Dict[str, World] = {}
        self._views: Dict[tuple, World] = {}
        self._pending_worlds: List[str] = []
        self._initialized = False
        self._time_ns = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step_size(self) -> float:
        return self._step_size

    def real_time_factor(self) -> float:
        return self._rtf

    def steps_per_run(self) -> int:
        return self._steps

    def initialized(self) -> bool:
        return self._initialized

    def running(self) -> bool:
        return False

    def pause(self) -> bool:
        return True

    def gui(self, verbosity: int = -1) -> bool:
        _warn("The B200 engine has no GUI")
        return False

    def insert_world_from_sdf(self, world_file: str = "", world_name: str = "") -> bool:
        if self._initialized:  # GazeboSimulator.cpp:390-433
            _err("Worlds must be inserted before the initialization")
            return False
        name = world_name
        if world_file:
            name = name or get_world_name_from_sdf(world_file)
            if not name:
                _err(f"Failed to read the world name from '{world_file}'")
                return False
        name = name or "default"  # utils.cpp:171-196 get_empty_world
        if name in self._pending_worlds:
            _err(f"A world named '{name}' was already inserted")
            return False
        self._pending_worlds.append(name)
        return True

    def insert_worlds_from_sdf(self, world_file: str, world_names: Sequence[str] = ()) -> bool:
        try:
            root = ET.parse(world_file).getroot()
        except Exception as e:  # noqa: BLE001
            _err(f"Failed to parse '{world_file}': {e}")
            return False
        found = [w.get("name") for w in root.iter("world")]
        names = list(world_names) if world_names else found
        if len(names) != len(found):
            _err("The number of world names does not match the number of worlds in the file")
            return False
        return all([self.insert_world_from_sdf("", n) for n in names])

    def initialize(self) -> bool:
        if self._initialized:
            return True
        if not (self._step_size > 0 and self._rtf > 0 and self._steps > 0):  # GazeboSimulator.cpp:578-600
            _err("Invalid simulator configuration (step size, real time factor and iterations must be positive)")
            return False
        if not self._pending_worlds:
            self._pending_worlds.append("default")
        for name in self._pending_worlds:
            world = World(self, name)
            # no GPU, no engine: this raises (there is no CPU fallback to degrade to)
            world._engine = b2sim.Simulator(self._num_envs, self._step_size, self._steps, self._dtype, self._device)
            self._worlds[name] = world
            _registry.add(name)
        self._initialized = True
        return True

    def run(self, paused: bool = False) -> bool:
        if not self._initialized:
            _err("The simulator was not initialized")
            return False
        if not paused:
            self._time_ns += self._steps * int(round(self._step_size * 1e9))
        for world in self._worlds.values():
            try:
                world._run(paused, self._time_ns)
            except b2sim.B2Error as e:
                _err(str(e))
                return False
        return True

    def close(self) -> bool:
        for name, world in list(self._worlds.items()):
            if world._engine is not None:
                world._engine.close()
                world._engine = None
            _registry.remove(name)
        self._worlds.clear()
        self._views.clear()
        self._initialized = False
        return True

    def world_names(self) -> tuple:
        return tuple(self._worlds) if self._initialized else tuple(self._pending_worlds)

    def _env_views(self, world: World) -> List[World]:
        return [w for (n, e), w in self._views.items() if n == world._name]

    def get_world(self, world_name: str = "", env: int = 0) -> World:
        if not self._initialized:
            raise RuntimeError("The simulator was not initialized")
        if not world_name:
            if len(self._worlds) != 1:
                raise RuntimeError("The simulator handles more than one world: a name is required")
            world_name = next(iter(self._worlds))
        if world_name not in self._worlds:
            raise RuntimeError(f"Failed to find world '{world_name}'")
        base = self._worlds[world_name]
        if env == 0:
            return base
        if not 0 <= env < self._num_envs:
            raise RuntimeError(f"env index {env} out of range")
        key = (world_name, env)
        if key not in self._views:
            self._views[key] = World(self, world_name, env, shared=base)
        return self._views[key]


# ---------------------------------------------------------------------------------------------------
# free functions (cpp/scenario/gazebo/include/scenario/gazebo/utils.h:75-240)
# ---------------------------------------------------------------------------------------------------
def get_empty_world() -> str:
    return ("<?xml version='1.0'?><sdf version='1.7'><world name='default'>"
            "<physics default='true' type='ignored'></physics></world></sdf>")


def _root_of(path_or_string: str):
    if os.path.isfile(path_or_string):
        return ET.parse(path_or_string).getroot()
    return ET.fromstring(path_or_string)


def sdf_string_valid(sdf_string: str) -> bool:
    try:
        root = ET.fromstring(sdf_string)
    except ET.ParseError:
        return False
    return root.tag == "sdf"


def get_sdf_string(file_name: str) -> str:
    try:
        with open(file_name, "r") as f:
            return f.read()
    except OSError:
        _err(f"Failed to read '{file_name}'")
        return ""


def get_world_name_from_sdf(file_name: str, world_index: int = 0) -> str:
    try:
        worlds = list(_root_of(file_name).iter("world"))
        return worlds[world_index].get("name", "")
    except Exception:  # noqa: BLE001
        return ""


def get_model_name_from_sdf(file_name: str, model_index: int = 0) -> str:
    try:
        root = _root_of(file_name)
        if root.tag == "robot":
            return root.get("name", "")
        return list(root.iter("model"))[model_index].get("name", "")
    except Exception:  # noqa: BLE001
        return ""


def find_sdf_file(file_name: str) -> str:
    for base in [""] + os.environ.get("IGN_GAZEBO_RESOURCE_PATH", "").split(":") + \
            os.environ.get("B2SIM_MODEL_PATH", "").split(":"):
        candidate = os.path.join(base, file_name) if base else file_name
        if os.path.isfile(candidate):
            return os.path.abspath(candidate)
    return ""


def get_model_file_from_fuel(uri: str, use_cache: bool = False) -> str:
    _err("Ignition Fuel needs network access, which the B200 engine does not provide")
    return ""


def get_random_string(length: int) -> str:
    return "".join(random.choice(string.ascii_letters + string.digits) for _ in range(int(length)))


def urdfstring_to_sdfstring(urdf_string: str) -> str:
    """The engine's loader reads URDF directly, so the 'conversion' is the identity."""
    return urdf_string


def urdffile_to_sdfstring(urdf_file: str) -> str:
    return get_sdf_string(urdf_file)


def _broadcast(v: Sequence[float], n: int) -> List[float]:
    v = [float(x) for x in v]
    if len(v) == n:
        return v
    if len(v) == 1:
        return v * n
    raise RuntimeError("Wrong input arguments")  # std::invalid_argument, utils.cpp:281-283


def _is_approx(a: List[float], b: List[float]) -> bool:
    # Eigen isApprox: ||a - b||^2 <= eps^2 * min(||a||^2, ||b||^2), eps = 1e-12
    diff = sum((x - y) ** 2 for x, y in zip(a, b))
    return diff <= 1e-24 * min(sum(x * x for x in a), sum(y * y for y in b))


def normalize(input: Sequence[float], low: Sequence[float], high: Sequence[float]) -> tuple:
    """utils.cpp:273-328: 2 (x - low) / (high - low) - 1, infinite results replaced by the input."""
    x = [float(v) for v in input]
    if not x:
        raise RuntimeError("Wrong input arguments")
    lo, hi = _broadcast(low, len(x)), _broadcast(high, len(x))
    if _is_approx(hi, lo):
        return tuple(x)
    out = []
    for v, a, b in zip(x, lo, hi):
        try:
            r = 2.0 * (v - a) / (b - a) - 1
        except ZeroDivisionError:
            r = math.inf
        out.append(v if math.isinf(r) or math.isnan(r) else r)
    return tuple(out)


def denormalize(input: Sequence[float], low: Sequence[float], high: Sequence[float]) -> tuple:
    """utils.cpp:330-376: (x + 1) (high - low) / 2 + low."""
    x = [float(v) for v in input]
    if not x:
        raise RuntimeError("Wrong input arguments")
    lo, hi = _broadcast(low, len(x)), _broadcast(high, len(x))
    if _is_approx(hi, lo):
        return tuple(x)
    return tuple((v + 1) * (b - a) / 2.0 + a for v, a, b in zip(x, lo, hi))
