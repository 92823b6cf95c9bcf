"""Throughput of the coupled world kernel (BASELINE config C5: Panda + table + cube with finger contacts, 4,096
envs/GPU), fp64, computed-torque controller at the physics rate. One env-step = one launch of k_world_coupled.
Phases: fingers open (no robot contact), grasp (8 finger contact points + table), lifted (finger contacts only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__; __graft_entry__.load_package()
import numpy as np, torch
import b2sim, gym_ignition_models
from b2sim import _lib
from test_coupled_gpu import CUBE_URDF, TABLE_SDF, KP, KD
from test_coupled_cpu import Q0, BASE

ZC = 1.487
for n in [int(a) for a in sys.argv[1:]] or (4096, 32768):
    sim = b2sim.Simulator(n, 0.001, 1)
    sim.insert_model_file(gym_ignition_models.get_model_file("ground_plane"))
    panda = sim.insert_model_file(gym_ignition_models.get_model_file("panda"), pose=list(BASE) + [1.0, 0, 0, 0], name="panda")
    sim.insert_model(TABLE_SDF, pose=[0.307, 0.0, ZC - 0.05, 1.0, 0, 0, 0], name="table")
    cube = sim.insert_model(CUBE_URDF, pose=[0.307, 0.0, ZC, 1.0, 0, 0, 0], name="cube")
    sim.tensor(panda, _lib.BUF_STATE)[:, :9] = torch.as_tensor(Q0, device="cuda")
    for j in (7, 8):
        sim.lib.b2sim_set_max_generalized_force(sim.handle, panda, j, 500.0)
    sim.set_controller_period(panda, 0.001)
    sim.set_computed_torque(panda, KP, KD)
    for j in range(9):
        sim.set_joint(panda, _lib.FIELD_POSITION_TARGET, -1, j, Q0[j])
        sim.set_joint(panda, _lib.FIELD_VELOCITY_TARGET, -1, j, 0.0)
        sim.set_joint(panda, _lib.FIELD_ACCELERATION_TARGET, -1, j, 0.0)
    pos_t = sim.tensor(panda, _lib.BUF_POS_TARGET)

    def timed(tag, steps):
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(steps):
            sim.run()
        s1.record(); torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / steps
        z = sim.tensor(cube, _lib.BUF_BASE_STATE)[:, 2]
        print(f"n={n} {tag}: {ms * 1e3:.1f} us per env-step launch -> {n / ms * 1e3:.3e} env-steps/s; cube z mean "
              f"{z.mean().item():.4f}; contacts env0 {len(sim.contacts(0))}", flush=True)

    for _ in range(20):
        sim.run()
    timed("open", 100)
    pos_t[:, 7:] = 0.0
    timed("grasp", 300)
    pos_t[:, 3] += 0.15   # elbow towards straight: the hand (and the grasped cube) rises by ~6.6 cm
    timed("lift", 300)
    sim.close()
