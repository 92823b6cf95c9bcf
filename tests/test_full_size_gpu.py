"""The BASELINE.json configurations at their full env counts on one B200, checked through properties that do not
depend on the size: a window of envs against the CPU oracle, invariance to how the envs are sharded (the Philox
streams are keyed by the global env index), and identities every row of the outputs must satisfy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


def test_pendulum_65536_envs_open_loop_rollout(torch, oracle, model_files):
    """BASELINE config 2 (Pendulum-Gazebo-v0, 65,536 envs, fp64): 200 steps in one launch."""
    import b2sim
    n, T, seed, window = 65536, 200, 2, 384
    env = b2sim.BatchedTaskEnv("Pendulum-Gazebo-v0", n, seed=seed, max_episode_steps=150)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    act = ((torch.rand(T, n, device="cuda", generator=gen, dtype=torch.float64) * 2 - 1) * 50.0).float().double().contiguous()
    obs, rew, done = env.trajectory(act)
    torch.cuda.synchronize()
    # every row: (cos, sin) on the unit circle, done <=> |dq| beyond the float32-rounded bound or the time limit
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert ((obs[..., 0] ** 2 + obs[..., 1] ** 2 - 1).abs() < 1e-12).all()
    assert done[149].all() or done[:150].any(dim=0).all()          # TimeLimit(150) reaches every env that did not end earlier
    over = obs[..., 2].abs() > float(np.float32(10.0))
    assert (done.bool() | ~over).all()
    # a window of envs somewhere in the middle against the oracle
    w0 = 40000
    _, model = oracle.load_urdf(model_files["pendulum"])
    ref_state = oracle.sample_reset_batch(1, seed, w0, window, 0)
    elapsed = np.zeros(window, np.int32)
    o_ref, r_ref, d_ref = oracle.rollout(model, 1, act[:, w0:w0 + window].cpu().numpy(), ref_state, elapsed, max_episode_steps=150,
                                         seed=seed, env_offset=w0, first_step=1)
    assert np.array_equal(done[:, w0:w0 + window].cpu().numpy(), d_ref)
    np.testing.assert_allclose(obs[:, w0:w0 + window].cpu().numpy(), o_ref, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(rew[:, w0:w0 + window].cpu().numpy(), r_ref, rtol=1e-9, atol=1e-11)
    env.close()


def test_cartpole_swingup_1m_envs_sharding_and_rows(torch, oracle, model_files):
    """BASELINE config 3 (CartPoleContinuousSwingup, 1,048,576 envs): one simulator against two half-size shards with
    their env offsets (bit-exact), a window against the oracle, per-row identities."""
    import b2sim
    env_id, n, T, seed = "CartPoleContinuousSwingup-Gazebo-v0", 1 << 20, 40, 4
    whole = b2sim.BatchedTaskEnv(env_id, n, seed=seed, max_episode_steps=25)
    halves = [b2sim.BatchedTaskEnv(env_id, n // 2, seed=seed, env_offset=k * (n // 2), max_episode_steps=25) for k in range(2)]
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    act = ((torch.rand(T, n, device="cuda", generator=gen, dtype=torch.float64) * 2 - 1) * 200.0).float().double().contiguous()
    w0, window = 777000, 256
    _, model = oracle.load_urdf(model_files["cartpole"])
    ref_state = oracle.sample_reset_batch(4, seed, w0, window, 0)
    assert np.array_equal(whole.state[w0:w0 + window].cpu().numpy(), ref_state)
    elapsed = np.zeros(window, np.int32)
    o_ref, r_ref, d_ref = oracle.rollout(model, 4, act[:, w0:w0 + window].cpu().numpy(), ref_state, elapsed, max_episode_steps=25,
                                         seed=seed, env_offset=w0, first_step=1)
    for t in range(T):
        before = whole.state.clone()
        obs, rew, done = whole.step(act[t])
        for k, h in enumerate(halves):
            h.step(act[t, k * (n // 2):(k + 1) * (n // 2)].contiguous())
        if t in (0, 24, T - 1):
            torch.cuda.synchronize()
            assert torch.isfinite(obs).all()
            keep = ~done.bool()
            # observation = [x, dx, q, dq] of the stepped state wherever the episode goes on
            st = whole.state
            assert torch.equal(obs[keep], st[keep][:, [0, 2, 1, 3]])
            # the cart moved by dx dt: semi-implicit Euler on every row
            assert ((obs[:, 0] - (before[:, 0] + obs[:, 1] * 1e-3)).abs() < 1e-12).all()
            # finished episodes restart inside the task's reset ranges (cartpole_continuous_swingup.py:145-146)
            if done.any():
                fresh = st[done.bool()]
                assert (fresh[:, [0, 2, 3]].abs() <= 0.05).all() and ((fresh[:, 1] - np.pi).abs() <= np.deg2rad(60) + 1e-12).all()
            np.testing.assert_allclose(obs[w0:w0 + window].cpu().numpy(), o_ref[t], rtol=1e-9, atol=1e-11)
            assert np.array_equal(done[w0:w0 + window].cpu().numpy(), d_ref[t])
    torch.cuda.synchronize()
    joined = torch.cat([h.state for h in halves])
    assert torch.equal(joined, whole.state), "two shards with env offsets must reproduce the single simulator bit for bit"
    assert torch.equal(torch.cat([h.elapsed for h in halves]), whole.elapsed)
    for e in [whole] + halves:
        e.close()


def test_panda_reach_16384_envs_rows(torch):
    """BASELINE config 4 (Panda PID + KinDyn observation, 16,384 envs): identities of every observation row."""
    import b2sim
    from b2sim.batched import PANDA_Q0
    n = 16384
    env = b2sim.BatchedTaskEnv("PandaReach-Gazebo-v0", n, seed=0)
    q0 = torch.tensor(PANDA_Q0, device="cuda", dtype=torch.float64)
    phase = torch.rand(n, 1, device="cuda", dtype=torch.float64) * 6.2831853
    tg = (q0 + 0.1 * torch.sin(phase)).contiguous()
    tg[:, 7:] = 0.02
    prev = None
    for t in range(30):
        obs, rew, done = env.step(tg)
        if t == 28:
            prev = obs.clone()
    torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and not done.any()
    assert torch.equal(obs[:, :18], env.state)                                   # q, dq
    assert ((obs[:, 21:25] ** 2).sum(dim=1) - 1).abs().max() < 1e-12             # unit quaternion
    J = obs[:, 25:].reshape(n, 6, 15)
    eye = torch.eye(3, device="cuda", dtype=torch.float64)
    assert torch.equal(J[:, :3, :3], eye.expand(n, 3, 3)) and torch.equal(J[:, 3:, 3:6], eye.expand(n, 3, 3))
    assert (J[:, 3:, :3] == 0).all() and (J[:, :, 13:] == 0).all()               # fingers are not on the chain
    goal = torch.tensor([0.5, 0.0, 0.5], device="cuda", dtype=torch.float64)
    assert (rew + (obs[:, 18:21] - goal).norm(dim=1)).abs().max() < 1e-12
    # the end effector moved by J dq dt (first order in dt)
    lin = torch.einsum("nij,nj->ni", J[:, :3, 6:], obs[:, 9:18])
    assert ((obs[:, 18:21] - prev[:, 18:21]) / 1e-3 - lin).abs().max() < 5e-3
    env.close()


def test_pick_scene_4096_envs(torch):
    """BASELINE config 5 (Panda + table + cube, 4,096 envs): close the gripper on the cube and lift it in every env."""
    import b2sim
    n = 4096
    sc = b2sim.PandaPickScene(n, seed=1)
    sc.step(30)
    sc.set_fingers(0.0)
    sc.step(250)
    z0 = sc.cube_state[:, 2].clone()
    sc.targets[:, 3] += 0.15       # elbow joint towards straight: the hand rises by ~6.6 cm
    sc.step(400)
    torch.cuda.synchronize()
    assert torch.isfinite(sc.state).all() and torch.isfinite(sc.cube_state).all()
    lifted = sc.cube_state[:, 2] - z0
    assert (lifted > 0.05).all(), lifted.min().item()                            # the grasp holds while lifting
    assert ((sc.cube_state[:, 3:7] ** 2).sum(dim=1) - 1).abs().max() < 1e-9      # unit quaternion after 680 steps
    fingers = sc.state[:, 7:9]
    assert (fingers > 0.02).all() and (fingers < 0.03).all()                     # pads stopped by the 5 cm cube
    sc.close()
