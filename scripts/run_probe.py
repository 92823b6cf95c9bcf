"""A/B of GazeboSimulator::run on a batch of Pandas under position PIDs (the generic scenario path, b2sim_run): one thread
per env (k_run_tree) against G lanes per env (k_run_tree_lanes). Each variant runs in its own process (B2_RUN_KERNEL is
read once); the parent compares the joint states after T runs and prints the time per run.

    python scripts/run_probe.py [n_envs] [steps_per_run]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(n, spr, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.load_package()
    import numpy as np
    import torch
    import b2sim
    import gym_ignition_models
    from b2sim import _lib as L
    from b2sim.batched import PANDA_Q0
    from gym_ignition_environments.models.panda import Panda  # noqa: F401  (gains live in the wrapper)
    sim = b2sim.Simulator(n, 0.001, spr)
    mid = sim.insert_model_file(gym_ignition_models.get_model_file("panda"))
    for j in range(9):
        sim.set_joint(mid, L.FIELD_POSITION_RESET, -1, j, PANDA_Q0[j])
    sim.run(paused=True)
    # gains of gym_ignition_environments/models/panda.py
    gains = [(50, 0, 20), (10000, 0, 500), (100, 0, 10), (1000, 0, 50), (100, 0, 10), (100, 0, 10), (10, 0.5, 0.1), (100, 0, 50),
             (100, 0, 50)]
    big = 1.7976931348623157e308
    sim.set_controller_period(mid, 0.001)  # the default (duration::max) computes the PID once and holds it
    if os.environ.get("PROBE_CONTROLLER") == "ct":  # ComputedTorqueFixedBase with the gains of examples/panda_pick_and_place.py
        sim.set_computed_torque(mid, [100.0] * 7 + [10000.0] * 2, [17.5] * 7 + [100.0] * 2)
        for j in range(9):
            sim.set_joint(mid, L.FIELD_VELOCITY_TARGET, -1, j, 0.0)
            sim.set_joint(mid, L.FIELD_ACCELERATION_TARGET, -1, j, 0.0)
            sim.set_joint(mid, L.FIELD_POSITION_TARGET, -1, j, PANDA_Q0[j])
    else:
        for j, (p, i, d) in enumerate(gains):
            sim.set_pid(mid, j, p, i, d, big, -big, big, -big, 0.0)
            sim.set_control_mode(mid, j, 5)
    tg = sim.tensor(mid, L.BUF_POS_TARGET)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    q0 = torch.tensor(PANDA_Q0, device="cuda", dtype=torch.float64)
    tg.copy_(q0 + 0.1 * torch.sin(torch.rand(n, 1, device="cuda", generator=gen, dtype=torch.float64) * 6.28))
    # fingers held mid-range, or (PROBE_FINGERS=limit) pressed against their lower limit: two joint-limit rows per env
    tg[:, 7:] = -0.01 if os.environ.get("PROBE_FINGERS") == "limit" else 0.02
    for _ in range(100):
        sim.run()
    torch.cuda.synchronize()
    np.savez(out, state=sim.tensor(mid, L.BUF_STATE).cpu().numpy()[:4096])
    best = 1e9
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            sim.run()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 100)
    print(f"{os.environ.get('B2_RUN_KERNEL', 'default'):7s} fingers={os.environ.get('PROBE_FINGERS', 'free')} controller={os.environ.get('PROBE_CONTROLLER', 'pid')} n={n} steps_per_run={spr}: {best * 1e3:8.1f} us/run  "
          f"{n * spr / best * 1e3:.3e} physics env-steps/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    import numpy as np
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    spr = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    outs = {}
    for variant in ("thread", "lanes"):
        out = os.path.join(ROOT, "gpurun_out", f"run_{variant}.npz")
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", str(n), str(spr), out],
                              env=dict(os.environ, B2_RUN_KERNEL=variant))
        outs[variant] = np.load(out)["state"]
    d = np.abs(outs["thread"] - outs["lanes"]) / (1e-9 + np.abs(outs["thread"]))
    print(f"state after 100 runs: max rel diff thread vs lanes = {d.max():.3e}")
