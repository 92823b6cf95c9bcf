"""Multi-GPU plumbing: env-index sharding and the optional end-of-rollout statistics all-reduce.

The step path has no collective (envs are independent worlds; the reference's own advice is one simulator per
robot, docs/sphinx/info/limitations.rst:19-20). ``torch.distributed`` is used for two things only: the
sum of a 4-value statistics vector at the end of a rollout, and barriers around timed regions.
"""
from typing import Tuple


def shard_range(global_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of global env indices owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(global_envs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class EpisodeStats:
    """Episode statistics [sum of returns, sum of lengths, finished episodes, non-finite rewards] of a rollout and their
    sum over ranks.

    ``EpisodeStats.from_env(env)`` is the product path: the fused step kernels accumulate the totals on the device
    (per-warp reduction + atomics into a striped vector, csrc/b2_kernels.cuh episode_stats_accumulate), nothing runs
    per step on the host, and ``all_reduce`` sums the 4 doubles over ranks with one NCCL all-reduce on the device
    tensor. The constructor with ``num_envs`` keeps the host-side accumulation from reward / done tensors
    (``update``), which the gloo tests on CPU use as the reference of the kernel-side accumulation."""

    def __init__(self, num_envs: int, device, dtype=None):
        import torch
        self._torch = torch
        self._env = None
        dtype = dtype or torch.float64
        self.ret = torch.zeros(num_envs, dtype=dtype, device=device)
        self.length = torch.zeros(num_envs, dtype=torch.int64, device=device)
        self.totals = torch.zeros(4, dtype=torch.float64, device=device)

    @classmethod
    def from_env(cls, env) -> "EpisodeStats":
        """Statistics accumulated by the kernels of a BatchedTaskEnv (enables them if needed)."""
        import torch
        self = cls.__new__(cls)
        self._torch = torch
        self._env = env
        env.enable_episode_stats(True)
        self._striped = env.sim.episode_stats_tensor(env.model)   # [stripes, 4] float64, zero-copy
        self.ret = env.sim.tensor(env.model, _EP_RETURN())
        self.length = env.elapsed
        return self

    @property
    def device_totals(self):
        """[4] float64 tensor on the device (a sum over the stripes; no host synchronisation)."""
        return self._striped.sum(dim=0) if self._env is not None else self.totals

    def update(self, reward, done) -> None:
        if self._env is not None:
            raise RuntimeError("kernel-side statistics are accumulated by env.step(); update() is the host-side path")
        torch = self._torch
        finite = torch.isfinite(reward)
        self.ret += torch.where(finite, reward, torch.zeros_like(reward)).to(self.ret.dtype)
        self.length += 1
        d = done.bool()
        self.totals[0] += self.ret[d].sum().double()
        self.totals[1] += self.length[d].sum().double()
        self.totals[2] += d.sum().double()
        self.totals[3] += (~finite).sum().double()
        self.ret[d] = 0
        self.length[d] = 0

    def all_reduce(self):
        """Sum the totals over all ranks (NCCL for CUDA tensors, gloo for CPU tensors). Returns a dict."""
        import torch.distributed as dist
        t = self.device_totals.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        s_ret, s_len, n, bad = (float(v) for v in t.tolist())
        return {"episodes": n, "mean_return": s_ret / n if n else float("nan"),
                "mean_length": s_len / n if n else float("nan"), "non_finite_rewards": bad}


def _EP_RETURN():
    from . import _lib
    return _lib.BUF_EP_RETURN
