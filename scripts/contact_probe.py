"""Throughput of the free-body contact kernel (BASELINE config C5 size: 4,096 envs/GPU): two stacked cubes on
the ground, fp64. One world step = one launch of k_world_free."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.load_package()
import numpy as np, torch
import b2sim, gym_ignition_models

CUBE = """<robot name="cube_robot"><link name="cube"><inertial><origin rpy="0 0 0" xyz="0 0 0"/><mass value="5.0"/>
<inertia ixx="0.0333333" ixy="0" ixz="0" iyy="0.0333333" iyz="0" izz="0.0333333"/></inertial>
<collision><geometry><box size="0.2 0.2 0.2"/></geometry><origin rpy="0 0 0" xyz="0 0 0"/></collision></link></robot>"""
for n in [int(a) for a in sys.argv[1:]] or (4096, 65536):
    sim = b2sim.Simulator(n, 0.001, 1)
    sim.insert_model_file(gym_ignition_models.get_model_file("ground_plane"))
    a = sim.insert_model(CUBE, pose=(0, 0, 0.15, 1, 0, 0, 0), name="a")
    b = sim.insert_model(CUBE, pose=(0.01, 0, 0.4, 1, 0, 0, 0), name="b")
    for _ in range(600):
        sim.run()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(200):
        sim.run()
    s1.record(); torch.cuda.synchronize()
    ms = s0.elapsed_time(s1) / 200
    z = sim.tensor(b, 14)[:, 2]
    print(f"n={n}: {ms * 1e3:.1f} us per world step -> {n / ms * 1e3:.3e} env-steps/s; upper cube z mean {z.mean().item():.4f}, "
          f"contacts env0 {len(sim.contacts(0))}")
    sim.close()
