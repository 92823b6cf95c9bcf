"""World-size-2 gloo tests of the multi-GPU host logic (runs on CPU): env-index sharding with the globally
keyed reset RNG reproduces the unsharded rollout bit for bit, and the statistics all-reduce sums over ranks.
The per-env arithmetic on this path is the oracle's (CPU), the sharding / collective logic is the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, T, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__
    __graft_entry__.load_package()
    import gym_ignition_models
    from b2sim.distributed import EpisodeStats, shard_range
    from oracle import oracle as O

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = shard_range(n_global, rank, world)
    n = stop - start
    task, seed = O.TASK_CARTPOLE_CONTINUOUS_SWINGUP, 9
    _, model = O.load_urdf(gym_ignition_models.get_model_file("cartpole"))
    actions = np.random.default_rng(0).uniform(-200, 200, (T, n_global))[:, start:stop].copy()
    state = O.sample_reset_batch(task, seed, start, n, 0)
    elapsed = np.zeros(n, np.int32)
    obs, rew, done = O.rollout(model, task, actions, state, elapsed, max_episode_steps=40, seed=seed, env_offset=start)
    stats = EpisodeStats(n, "cpu")
    for t in range(T):
        stats.update(torch.as_tensor(rew[t]), torch.as_tensor(done[t]))
    summary = stats.all_reduce()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=state, done=done, rew=rew, start=start, stop=stop,
             episodes=summary["episodes"], mean_length=summary["mean_length"])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_the_env_indices():
    from b2sim.distributed import shard_range
    for n, w in ((1048576, 8), (10, 3), (7, 8), (5, 1)):
        blocks = [shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def test_two_rank_sharded_rollout_equals_unsharded(tmp_path, oracle, model_files):
    n_global, T, world = 96, 120, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_global, T, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # unsharded run
    task, seed = oracle.TASK_CARTPOLE_CONTINUOUS_SWINGUP, 9
    _, model = oracle.load_urdf(model_files["cartpole"])
    actions = np.random.default_rng(0).uniform(-200, 200, (T, n_global))
    state = oracle.sample_reset_batch(task, seed, 0, n_global, 0)
    elapsed = np.zeros(n_global, np.int32)
    _, rew, done = oracle.rollout(model, task, actions, state, elapsed, max_episode_steps=40, seed=seed)
    assert np.array_equal(np.concatenate([p["state"] for p in parts]), state)
    assert np.array_equal(np.concatenate([p["done"] for p in parts], axis=1), done)
    assert np.array_equal(np.concatenate([p["rew"] for p in parts], axis=1), rew)
    # the all-reduce saw every rank's episodes
    assert parts[0]["episodes"] == parts[1]["episodes"] == done.sum()
    assert parts[0]["mean_length"] == pytest.approx(parts[1]["mean_length"])
