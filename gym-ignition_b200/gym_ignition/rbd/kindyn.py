"""KinDyn queries on the B200 engine (replaces the iDynTree wrapper of
python/gym_ignition/rbd/idyntree/kindyncomputations.py for fixed-base models).

The same method names are kept where the quantity is provided: ``set_robot_state``,
``set_robot_state_from_model``, ``get_joint_positions`` / ``velocities``, ``get_world_transform``,
``get_relative_transform``, ``get_frame_jacobian`` (6 x (6+n), MIXED representation, linear rows first,
kindyncomputations.py:367-377), ``get_mass_matrix``, ``get_bias_forces``, ``get_com_position`` / ``velocity``,
``get_momentum``, ``get_centroidal_momentum``, ``get_average_velocity``, ``get_centroidal_average_velocity``, the four
momentum / average-velocity Jacobians, ``get_frame_bias_acc`` and ``get_com_bias_acc``. The batched variants
(``*_batch``) return CUDA tensors for every env of the query simulator.

Fixed base: the (6+n) quantities of iDynTree reduce to their joint blocks plus, for the Jacobian, the analytic
base block [[1, -S(p_frame - p_base)], [0, 1]]. ``get_mass_matrix`` / ``get_bias_forces`` return the joint
blocks (n x n, n), which is what the reference's fixed-base controller consumes
(cpp/scenario/controllers/src/ComputedTorqueFixedBase.cpp:312-327).
"""
from typing import List, Optional

import b2sim
import numpy as np
from b2sim import _lib as _b2

from . import conversions


class KinDynComputations:
    def __init__(self, model_file: str, considered_joints: List[str] = None,
                 world_gravity: np.ndarray = np.array([0, 0, -9.806]), num_envs: int = 1, dtype: str = "float64",
                 device: int = 0):
        import torch
        self._torch = torch
        self.world_gravity = np.array(world_gravity, float)
        self.sim = b2sim.Simulator(num_envs, 0.001, 1, dtype, device)
        self.sim.set_gravity(self.world_gravity)
        self.model = self.sim.insert_model_file(model_file)
        info = self.sim.info(self.model)
        self._joint_names = list(info.joint_names)
        self._link_names = list(info.link_names)
        self.dofs = info.dofs
        self._considered_joints = list(considered_joints) if considered_joints is not None else self._joint_names
        unknown = set(self._considered_joints) - set(self._joint_names)
        if unknown:
            raise ValueError(f"unknown joints {sorted(unknown)}")
        self._perm = [self._joint_names.index(n) for n in self._considered_joints]
        self._state = self.sim.tensor(self.model, _b2.BUF_STATE)
        self._tdt = self._state.dtype
        self._dev = self._state.device
        self._base_H = np.eye(4)
        self.num_envs = num_envs

    def joint_serialization(self) -> List[str]:
        return self._considered_joints

    def get_floating_base(self) -> str:
        return self._link_names[0]

    # ---- state ----
    def set_robot_state(self, s: np.ndarray, ds: np.ndarray, world_H_base: np.ndarray = np.eye(4),
                        base_velocity: np.ndarray = np.zeros(6), world_gravity: np.ndarray = None) -> None:
        s, ds = np.asarray(s, float).ravel(), np.asarray(ds, float).ravel()
        if s.size != len(self._perm) or ds.size != len(self._perm):
            raise ValueError("wrong size of the joint state")
        if np.asarray(world_H_base).shape != (4, 4):
            raise ValueError(world_H_base)
        if not np.allclose(world_H_base, self._base_H):
            raise ValueError("the fixed base cannot be moved after the model was loaded")
        row = np.zeros(2 * self.dofs)
        row[self._perm] = s
        row[[self.dofs + j for j in self._perm]] = ds
        self._state.copy_(self._torch.as_tensor(row, dtype=self._tdt, device=self._dev).expand_as(self._state))

    def set_robot_state_batch(self, q, dq) -> None:
        """q, dq: CUDA tensors [num_envs, dofs] in the model's joint order."""
        self._state[:, :self.dofs].copy_(q)
        self._state[:, self.dofs:].copy_(dq)

    def set_robot_state_from_model(self, model, world_gravity: np.ndarray = None) -> None:
        s = np.array(model.joint_positions(self._considered_joints))
        ds = np.array(model.joint_velocities(self._considered_joints))
        self.set_robot_state(s, ds)

    def get_joint_positions(self) -> np.ndarray:
        return self._state[0, :self.dofs].double().cpu().numpy()[self._perm]

    def get_joint_velocities(self) -> np.ndarray:
        return self._state[0, self.dofs:].double().cpu().numpy()[self._perm]

    def get_model_velocity(self) -> np.ndarray:
        return np.concatenate([np.zeros(6), self.get_joint_velocities()])

    # ---- kinematics ----
    def _link(self, frame_name: str) -> int:
        if frame_name not in self._link_names:
            raise RuntimeError(f"Frame '{frame_name}' does not exist")
        return self._link_names.index(frame_name)

    def world_transform_batch(self, frame_name: str):
        """[num_envs, 7] xyz + quaternion wxyz of ``frame_name``."""
        l = self._link(frame_name)
        self.sim.update_kinematics(self.model)
        return self.sim.tensor(self.model, _b2.BUF_LINK_POSE).view(self.num_envs, -1, 7)[:, l]

    def get_world_transform(self, frame_name: str) -> np.ndarray:
        pose = self.world_transform_batch(frame_name)[0].double().cpu().numpy()
        return conversions.Transform.from_position_and_quaternion(pose[:3], pose[3:])

    def get_world_base_transform(self) -> np.ndarray:
        return self.get_world_transform(self.get_floating_base())

    def get_relative_transform(self, ref_frame_name: str, frame_name: str) -> np.ndarray:
        return np.linalg.inv(self.get_world_transform(ref_frame_name)) @ self.get_world_transform(frame_name)

    def frame_jacobian_batch(self, frame_name: str):
        """[num_envs, 6, dofs]: joint columns of the MIXED frame Jacobian (model joint order)."""
        J = self._torch.empty((self.num_envs, 6 * self.dofs), dtype=self._tdt, device=self._dev)
        self.sim.kindyn(self.model, self._link(frame_name), None, None, J)
        return J.view(self.num_envs, 6, self.dofs)

    def get_frame_jacobian(self, frame_name: str) -> np.ndarray:
        Jj = self.frame_jacobian_batch(frame_name)[0].double().cpu().numpy()[:, self._perm]
        p_frame = self.get_world_transform(frame_name)[:3, 3]
        p_base = self.get_world_base_transform()[:3, 3]
        r = p_frame - p_base
        S = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        Jb = np.block([[np.eye(3), -S], [np.zeros((3, 3)), np.eye(3)]])
        return np.hstack([Jb, Jj])

    # ---- dynamics ----
    def mass_matrix_batch(self):
        M = self._torch.empty((self.num_envs, self.dofs * self.dofs), dtype=self._tdt, device=self._dev)
        self.sim.kindyn(self.model, 0, M, None, None)
        return M.view(self.num_envs, self.dofs, self.dofs)

    def bias_forces_batch(self):
        h = self._torch.empty((self.num_envs, self.dofs), dtype=self._tdt, device=self._dev)
        self.sim.kindyn(self.model, 0, None, h, None)
        return h

    def get_mass_matrix(self) -> np.ndarray:
        M = self.mass_matrix_batch()[0].double().cpu().numpy()
        return M[np.ix_(self._perm, self._perm)]

    def get_bias_forces(self) -> np.ndarray:
        return self.bias_forces_batch()[0].double().cpu().numpy()[self._perm]

    def get_generalized_gravity_forces(self) -> np.ndarray:
        saved = self._state[:, self.dofs:].clone()
        self._state[:, self.dofs:].zero_()
        g = self.get_bias_forces()
        self._state[:, self.dofs:].copy_(saved)
        return g

    # ---- centre of mass and momentum (kindyncomputations.py:305-342) ----
    def centroidal_batch(self):
        """(com [N,3], com velocity [N,3], momentum [N,12]: about the world origin then about the centre of mass,
        centre-of-mass Jacobian [N,3,dofs]) for every env, as CUDA tensors."""
        mk = lambda *shape: self._torch.empty(shape, dtype=self._tdt, device=self._dev)
        com, vel, mom, jac = mk(self.num_envs, 3), mk(self.num_envs, 3), mk(self.num_envs, 12), mk(self.num_envs, 3 * self.dofs)
        self.sim.centroidal(self.model, com, vel, mom, jac)
        return com, vel, mom, jac.view(self.num_envs, 3, self.dofs)

    def get_com_position(self) -> np.ndarray:
        return self.centroidal_batch()[0][0].double().cpu().numpy()

    def get_com_velocity(self) -> np.ndarray:
        """MIXED representation (world orientation); the base of these models does not move."""
        return self.centroidal_batch()[1][0].double().cpu().numpy()

    def get_momentum(self):
        """Linear and angular momentum in the frame iDynTree's MIXED representation expresses them in (world
        orientation, origin at the base): the kernel's moment about the world origin is moved to the base origin."""
        mom = self.centroidal_batch()[2][0].double().cpu().numpy()
        p_base = self.get_world_base_transform()[:3, 3]
        return mom[0:3], mom[3:6] - np.cross(p_base, mom[0:3])

    def get_centroidal_momentum(self):
        mom = self.centroidal_batch()[2][0].double().cpu().numpy()
        return mom[6:9], mom[9:12]

    def get_com_jacobian(self) -> np.ndarray:
        """3 x (6 + n): fixed base, so the base block is [1, -S(com - p_base)] like the frame Jacobians."""
        com, _, _, jac = self.centroidal_batch()
        r = com[0].double().cpu().numpy() - self.get_world_base_transform()[:3, 3]
        S = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
        return np.hstack([np.eye(3), -S, jac[0].double().cpu().numpy()[:, self._perm]])

    # ---- momentum Jacobians and average velocities (kindyncomputations.py:351-363, 379-427) ----
    # nu = [base linear velocity, base angular velocity (world orientation, MIXED), joint velocities]; the momentum is
    # h = J nu with the base block equal to the locked 6D inertia about the base origin. The base of these models is
    # fixed (nu[:6] = 0); the base blocks are returned because the reference's matrices are 6 x (6 + n).
    def momentum_jacobian_batch(self):
        """(joint block of the momentum Jacobian [N, 6, dofs], locked inertia about the base origin [N, 10]) as CUDA
        tensors, world orientation, linear rows first (include/b2sim.h b2sim_momentum_jacobian)."""
        J = self._torch.empty((self.num_envs, 6 * self.dofs), dtype=self._tdt, device=self._dev)
        locked = self._torch.empty((self.num_envs, 10), dtype=self._tdt, device=self._dev)
        self.sim.momentum_jacobian(self.model, J, locked)
        return J.view(self.num_envs, 6, self.dofs), locked

    @staticmethod
    def _skew(r):
        return np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])

    def _momentum_terms(self):
        J, locked = self.momentum_jacobian_batch()
        Jj = J[0].double().cpu().numpy()[:, self._perm]
        lk = locked[0].double().cpu().numpy()
        A = np.array([[lk[0], lk[1], lk[2]], [lk[1], lk[3], lk[4]], [lk[2], lk[4], lk[5]]])
        mc, mass = lk[6:9], lk[9]
        S = self._skew(mc)
        locked6 = np.block([[mass * np.eye(3), -S], [S, A]])
        return Jj, locked6, mc / mass, mass

    def get_linear_angular_momentum_jacobian(self) -> np.ndarray:
        """6 x (6 + n): momentum about the base origin, world orientation (MIXED)."""
        Jj, locked6, _, _ = self._momentum_terms()
        return np.hstack([locked6, Jj])

    def get_centroidal_total_momentum_jacobian(self) -> np.ndarray:
        """6 x (6 + n): the same momentum taken about the centre of mass."""
        Jj, locked6, c, _ = self._momentum_terms()
        X = np.block([[np.eye(3), np.zeros((3, 3))], [-self._skew(c), np.eye(3)]])
        return X @ np.hstack([locked6, Jj])

    def get_average_velocity_jacobian(self) -> np.ndarray:
        """6 x (6 + n): locked inertia^-1 times the momentum Jacobian (base block = identity)."""
        Jj, locked6, _, _ = self._momentum_terms()
        return np.linalg.solve(locked6, np.hstack([locked6, Jj]))

    def get_centroidal_average_velocity_jacobian(self) -> np.ndarray:
        Jj, locked6, c, mass = self._momentum_terms()
        X = np.block([[np.eye(3), np.zeros((3, 3))], [-self._skew(c), np.eye(3)]])
        lockedG = X @ locked6 @ X.T  # locked inertia about the centre of mass: blockdiag(m 1, I_G)
        return np.linalg.solve(lockedG, X @ np.hstack([locked6, Jj]))

    def get_average_velocity(self) -> np.ndarray:
        return self.get_average_velocity_jacobian() @ self.get_model_velocity()

    def get_centroidal_average_velocity(self) -> np.ndarray:
        return self.get_centroidal_average_velocity_jacobian() @ self.get_model_velocity()

    def get_com_bias_acc(self) -> np.ndarray:
        """dJ_com nu: acceleration of the centre of mass at zero joint acceleration (kindyncomputations.py:423-426),
        the mass-weighted classical accelerations of the link centres of mass."""
        tb = self.sim.info(self.model).tables()
        self.sim.tensor(self.model, _b2.BUF_ACCELERATION).zero_()
        twist = self._torch.empty((self.num_envs, 6), dtype=self._tdt, device=self._dev)
        acc = self._torch.empty((self.num_envs, 6), dtype=self._tdt, device=self._dev)
        total, out = float(tb["total_mass"]), np.zeros(3)
        for l, name in enumerate(self._link_names):
            ml = float(tb["link_mass"][l])
            if ml == 0.0 or tb["link_body"][l] < 0:
                continue
            self.sim.link_motion(self.model, l, twist, acc)
            v, a = twist[0].double().cpu().numpy(), acc[0].double().cpu().numpy()
            r = self.get_world_transform(name)[:3, :3] @ tb["link_com"][l]
            w = v[3:]
            out += ml * (a[:3] + np.cross(a[3:], r) + np.cross(w, np.cross(w, r)))
        return out / total

    def get_frame_bias_acc(self, frame_name: str) -> np.ndarray:
        """dJ nu of the frame (MIXED): its acceleration [linear, angular] at zero joint acceleration."""
        acc = self._torch.empty((self.num_envs, 6), dtype=self._tdt, device=self._dev)
        self.sim.tensor(self.model, _b2.BUF_ACCELERATION).zero_()
        self.sim.link_motion(self.model, self._link(frame_name), None, acc)
        return acc[0].double().cpu().numpy()

    def close(self) -> None:
        self._state = None
        self.sim.close()
