"""Throughput of the single-launch trajectory kernel against one launch per step, at small and large env counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__; __graft_entry__.load_package()
import torch, b2sim

def timed(fn, reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for env_id, amp, traj_bytes in (("Pendulum-Gazebo-v0", 50.0, 8 + 24 + 8 + 1), ("CartPoleContinuousSwingup-Gazebo-v0", 200.0, 8 + 32 + 8 + 1)):
    for n in (16384, 65536, 262144, 1048576):
        T = 250 if n <= 262144 else 64
        env = b2sim.BatchedTaskEnv(env_id, n, seed=0)
        act = ((torch.rand(T, n, device="cuda", dtype=torch.float64) * 2 - 1) * amp).contiguous()
        bufs = (torch.empty((T, n, env.nobs), dtype=torch.float64, device="cuda"),
                torch.empty((T, n), dtype=torch.float64, device="cuda"), torch.empty((T, n), dtype=torch.uint8, device="cuda"))
        for _ in range(2):
            env.rollout(act); env.trajectory(act, out=bufs); env.trajectory(act, record=False)
        ms_loop = timed(lambda: env.rollout(act), 4) / T
        ms_traj = timed(lambda: env.trajectory(act, out=bufs), 4) / T
        ms_last = timed(lambda: env.trajectory(act, record=False), 4) / T
        print(f"{env_id} n={n}: per-step launches {ms_loop * 1e3:.2f} us/step ({n / ms_loop * 1e3:.3e}/s) | one launch, trajectory written "
              f"{ms_traj * 1e3:.2f} us/step ({n / ms_traj * 1e3:.3e}/s, {n * traj_bytes / ms_traj / 1e6:.0f} GB/s) | one launch, last step only "
              f"{ms_last * 1e3:.2f} us/step ({n / ms_last * 1e3:.3e}/s)", flush=True)
        env.close()
