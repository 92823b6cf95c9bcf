"""Batched pick scene (BASELINE.json config 5): N worlds, each with a Panda on a 1 m pedestal, a table and a
5 cm cube between the open finger pads - the layout of examples/panda_pick_and_place.py:213-251 with the Fuel
meshes (table, wood cube) replaced by primitives, since Fuel needs the network. The arm runs the example's
ComputedTorqueFixedBase controller (:22-46) at the physics rate; finger targets open / close the gripper.

One env-step = one `GazeboSimulator::run`: controller, articulated-body dynamics, contact generation and ONE
constraint solve over joint limits and contacts (csrc/b2_kernels.cuh: k_coupled_dynamics | k_coupled_rows |
k_coupled_minv side by side, then k_pgs_solve and k_world_finish).
"""
import numpy as np

from . import _lib
from .batched import PANDA_Q0
from .engine import Simulator

CUBE_EDGE, CUBE_MASS = 0.05, 0.1
_I = CUBE_MASS / 12 * 2 * CUBE_EDGE ** 2
CUBE_URDF = f"""<robot name="cube"><link name="cube">
  <inertial><origin rpy="0 0 0" xyz="0 0 0"/><mass value="{CUBE_MASS}"/>
    <inertia ixx="{_I}" ixy="0" ixz="0" iyy="{_I}" iyz="0" izz="{_I}"/></inertial>
  <collision><geometry><box size="{CUBE_EDGE} {CUBE_EDGE} {CUBE_EDGE}"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/></collision>
</link></robot>"""
TABLE_SDF = """<?xml version="1.0"?>
<sdf version="1.7"><model name="table"><static>true</static><link name="top">
  <collision name="c"><geometry><box><size>0.4 0.4 0.05</size></box></geometry></collision>
</link></model></sdf>"""
KP = [100.0] * 7 + [10000.0] * 2   # examples/panda_pick_and_place.py:34-40
KD = [17.5] * 7 + [100.0] * 2
#: Algorithmic HBM bytes per env-step of THIS scene (it steps worlds through GazeboSimulator::run and has no Task, so the
#: observation-side bytes of SURVEY.md 8d's 2026 B estimate for config 5 are not moved), fp64, per-env tensors that must
#: cross HBM once per step: read q, dq (144) + position / velocity / acceleration references (216) + the controller's
#: held torque (72) + reset mask (4) + cube state (104) = 540; write q, dq (144) + joint accelerations (72) + held torque
#: (72) + cube state (104) + cube acceleration (48) + contact count (4) + the records of the grasp phase's 9 reported
#: contacts (9 x (16 + 80) = 864) = 1308. The solver's rows, M^-1 and impulses are intermediates (L2-resident at 4,096 envs).
ALGORITHMIC_BYTES = {"float64": 1848, "float32": 952}


class PandaPickScene:
    def __init__(self, num_envs: int, dtype: str = "float64", device: int = 0, seed: int = 0,
                 base=(0.0, 0.0, 1.0), cube_xy_jitter=(0.004, 0.01), steps_per_run: int = 1):
        import gym_ignition_models
        import torch
        self.torch = torch
        self.num_envs = num_envs
        self.sim = Simulator(num_envs, 0.001, steps_per_run, dtype, device)
        self.ground = self.sim.insert_model_file(gym_ignition_models.get_model_file("ground_plane"))
        self.panda = self.sim.insert_model_file(gym_ignition_models.get_model_file("panda"), pose=list(base) + [1.0, 0, 0, 0],
                                                name="panda")
        # the end effector of the initial configuration hovers at (0.307, 0, base_z + 0.487): the cube sits there
        self.cube_centre = np.array([base[0] + 0.307, base[1], base[2] + 0.487])
        table_centre = self.cube_centre - np.array([0, 0, CUBE_EDGE / 2 + 0.025])
        self.table = self.sim.insert_model(TABLE_SDF, pose=list(table_centre) + [1.0, 0, 0, 0], name="table")
        self.cube = self.sim.insert_model(CUBE_URDF, pose=list(self.cube_centre) + [1.0, 0, 0, 0], name="cube")
        self.q0 = np.array(PANDA_Q0, float)
        self.q0[7:] = 0.04
        self.state = self.sim.tensor(self.panda, _lib.BUF_STATE)
        self.cube_state = self.sim.tensor(self.cube, _lib.BUF_BASE_STATE)
        tdt = self.state.dtype
        self.state[:, :9] = torch.as_tensor(self.q0, device=self.state.device, dtype=tdt)
        gen = torch.Generator(device=self.state.device)
        gen.manual_seed(seed)
        jitter = (torch.rand(num_envs, 2, generator=gen, device=self.state.device, dtype=tdt) * 2 - 1) * \
            torch.as_tensor(cube_xy_jitter, device=self.state.device, dtype=tdt)
        self.cube_state[:, :2] += jitter
        for j in (7, 8):  # panda_pick_and_place.py:28-32
            _lib.check(self.sim.lib.b2sim_set_max_generalized_force(self.sim.handle, self.panda, j, 500.0))
        self.sim.set_controller_period(self.panda, 0.001)
        self.sim.set_computed_torque(self.panda, KP, KD)
        for j in range(9):
            self.sim.set_joint(self.panda, _lib.FIELD_POSITION_TARGET, -1, j, float(self.q0[j]))
            self.sim.set_joint(self.panda, _lib.FIELD_VELOCITY_TARGET, -1, j, 0.0)
            self.sim.set_joint(self.panda, _lib.FIELD_ACCELERATION_TARGET, -1, j, 0.0)
        self.targets = self.sim.tensor(self.panda, _lib.BUF_POS_TARGET)
        self.bytes_per_env_step = ALGORITHMIC_BYTES[dtype]

    def set_fingers(self, opening: float) -> None:
        """Position target of both finger joints for every env (0 = closed, 0.04 = open)."""
        self.targets[:, 7:] = opening

    def step(self, count: int = 1) -> None:
        for _ in range(count):
            self.sim.run()

    def close(self) -> None:
        self.state = self.cube_state = self.targets = None
        self.sim.close()
