"""A/B of the fused chain step kernels at the headline size: k_task_chain (one env per thread, one pass) against
k_task_chain_stream (grid-stride loop with the next env's loads in flight). Each variant runs in its own process
(the choice is read once from B2_CHAIN_KERNEL / B2_CHAIN_BLOCKS_PER_SM).

    python scripts/chain_probe.py [n_envs]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(n):
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.load_package()
    import torch
    import b2sim
    env = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=0)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    act4 = ((torch.rand(4, n, device="cuda", generator=gen, dtype=torch.float64) * 2 - 1) * 200.0).contiguous()
    for _ in range(10):
        env.rollout(act4)
    torch.cuda.synchronize()
    ts = []
    for rep in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(25):
            env.rollout(act4)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 100)
    ts.sort()
    gbs = 114 * n / (ts[0] * 1e-3) / 1e9
    print(f"{os.environ.get('B2_CHAIN_KERNEL', 'auto'):6s} blocks/SM={os.environ.get('B2_CHAIN_BLOCKS_PER_SM', '-'):2s} n={n}: "
          f"best {ts[0] * 1e3:7.1f} us  median {ts[3] * 1e3:7.1f} us  {gbs:7.0f} GB/s algorithmic ({gbs / 6547.2:.3f} of 6547)",
          flush=True)
    print("checksum", float(env.state.sum().item()), int(env.elapsed.sum().item()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1048576
    for variant, blocks in (("plain", None), ("stream", "2"), ("stream", "3"), ("stream", "4"), ("stream", "5"), ("stream", "6"),
                            ("stream", "8")):
        envv = dict(os.environ, B2_CHAIN_KERNEL=variant)
        if blocks:
            envv["B2_CHAIN_BLOCKS_PER_SM"] = blocks
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", str(n)], env=envv)
