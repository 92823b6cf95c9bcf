"""Times the generic tree path for the Panda PID + KinDyn workload (BASELINE config C4) kernel by kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.load_package()
import numpy as np, torch
import b2sim, gym_ignition_models
from b2sim import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dtype = sys.argv[2] if len(sys.argv) > 2 else "float64"
sim = b2sim.Simulator(n, 0.001, 1, dtype)
mid = sim.insert_model_file(gym_ignition_models.get_model_file("panda"))
q0 = [0, -0.785, 0, -2.356, 0, 1.571, 0.785, float(os.environ.get("FINGER", "0")), float(os.environ.get("FINGER", "0"))]
gains = [(50, 0, 20), (10000, 0, 500), (100, 0, 10), (1000, 0, 50), (100, 0, 10), (100, 0, 10), (10, 0.5, 0.1), (100, 0, 50), (100, 0, 50)]
for j in range(9):
    sim.set_joint(mid, L.FIELD_POSITION_RESET, -1, j, q0[j])
sim.run(paused=True)
sim.set_controller_period(mid, 0.001)
M = np.finfo(np.float64).max
for j, (p, i, d) in enumerate(gains):
    sim.set_pid(mid, j, p, i, d, M, -M, M, -M, 0.0)
    sim.set_control_mode(mid, j, 5)
tdt = torch.float64 if dtype == "float64" else torch.float32
J = torch.empty((n, 54), dtype=tdt, device="cuda")
ee = sim.info(mid).link_names.index("end_effector_frame")
pt = sim.tensor(mid, L.BUF_POS_TARGET)
base = pt.clone()
phase = torch.rand(n, 1, device="cuda", dtype=tdt) * 6.28

def step(k):
    pt.copy_(base + 0.1 * torch.sin(2 * np.pi * 0.33 * k * 0.001 + phase))
    sim.run()
    sim.kindyn(mid, ee, None, None, J)
    sim.update_kinematics(mid)

for k in range(20):
    step(k)
torch.cuda.synchronize()
for name, fn in (("run(PID+ABA)", lambda: sim.run()), ("kindyn(J)", lambda: sim.kindyn(mid, ee, None, None, J)),
                 ("kinematics", lambda: sim.update_kinematics(mid))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:14s} {a.elapsed_time(b) / 50 * 1e3:9.1f} us/launch  n={n} {dtype}")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for k in range(100):
    step(k)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 100
print(f"full step {ms * 1e3:.1f} us -> {n / ms * 1e3:.3e} env-steps/s")
print("q err deg", np.rad2deg((sim.tensor(mid, 0)[:, :9] - pt).abs().max().item()))
