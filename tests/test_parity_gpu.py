"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): done/reset masks and env indexing bit-exact; joint positions and
velocities within 1e-9 relative per step in fp64; within 1e-6 after 1000 contact-free steps.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TASKS = [
    ("Pendulum-Gazebo-v0", 1, "pendulum", 50.0),
    ("CartPoleDiscreteBalancing-Gazebo-v0", 2, "cartpole", None),
    ("CartPoleContinuousBalancing-Gazebo-v0", 3, "cartpole", 50.0),
    ("CartPoleContinuousSwingup-Gazebo-v0", 4, "cartpole", 200.0),
]


def make_actions(rng, T, n, amp):
    if amp is None:  # Discrete(2)
        return rng.integers(0, 2, (T, n)).astype(np.float64)
    return rng.uniform(-amp, amp, (T, n)).astype(np.float32).astype(np.float64)


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


@pytest.mark.parametrize("env_id,task,model_name,amp", TASKS)
def test_fused_rollout_matches_oracle(env_id, task, model_name, amp, torch, oracle, model_files):
    import b2sim
    n, T, seed, offset = 640, 400, 11, 1000
    env = b2sim.BatchedTaskEnv(env_id, n, seed=seed, env_offset=offset, max_episode_steps=150)
    _, model = oracle.load_urdf(model_files[model_name])
    ref_state = oracle.sample_reset_batch(task, seed, offset, n, 0)
    assert np.array_equal(env.state.cpu().numpy(), ref_state), "initial reset states must be bit-exact"
    rng = np.random.default_rng(5)
    actions = make_actions(rng, T, n, amp)
    elapsed = np.zeros(n, np.int32)
    o_ref, r_ref, d_ref = oracle.rollout(model, task, actions, ref_state, elapsed, max_episode_steps=150, seed=seed,
                                         env_offset=offset, first_step=1)
    a_dev = torch.as_tensor(actions, device="cuda")
    obs = torch.empty((T, n, env.nobs), dtype=torch.float64, device="cuda")
    rew = torch.empty((T, n), dtype=torch.float64, device="cuda")
    done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    for t in range(T):
        o, r, d = env.step(a_dev[t])
        obs[t].copy_(o); rew[t].copy_(r); done[t].copy_(d)
    torch.cuda.synchronize()
    assert d_ref.sum() > 0, "the rollout must exercise termination and auto-reset"
    assert np.array_equal(done.cpu().numpy(), d_ref), "done masks must be bit-exact"
    np.testing.assert_allclose(obs.cpu().numpy(), o_ref, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(rew.cpu().numpy(), r_ref, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(env.state.cpu().numpy(), ref_state, rtol=1e-9, atol=1e-11)
    assert np.array_equal(env.elapsed.cpu().numpy().astype(np.int32), elapsed)
    env.close()


@pytest.mark.parametrize("env_id,task,model_name,amp", [TASKS[0], TASKS[3]])
def test_one_step_from_identical_states_1e9(env_id, task, model_name, amp, torch, oracle, model_files):
    """Per-step bound: start both sides from the same random states, one step, 1e-9 relative."""
    import b2sim
    n = 4096
    env = b2sim.BatchedTaskEnv(env_id, n, seed=1)
    _, model = oracle.load_urdf(model_files[model_name])
    rng = np.random.default_rng(2)
    nq = oracle.task_nq(task)
    state = np.concatenate([rng.uniform(-2, 2, (n, nq)), rng.uniform(-5, 5, (n, nq))], axis=1)
    env.state.copy_(torch.as_tensor(state, device="cuda"))
    actions = make_actions(rng, 1, n, amp)
    ref = state.copy()
    elapsed = np.zeros(n, np.int32)
    oracle.rollout(model, task, actions, ref, elapsed, max_episode_steps=5000, seed=1)
    env.step(torch.as_tensor(actions[0], device="cuda"))
    torch.cuda.synchronize()
    got = env.state.cpu().numpy()
    keep = elapsed == 1  # envs that did not terminate (terminated ones were re-sampled, checked elsewhere)
    assert keep.sum() > n // 2
    np.testing.assert_allclose(got[keep], ref[keep], rtol=1e-9, atol=1e-13)
    env.close()


def test_contact_free_1000_steps_1e6(torch, oracle, model_files):
    """1e-6 after 1000 steps on contact-free cartpole and pendulum (no terminations in between)."""
    import b2sim
    for env_id, task, model_name, amp in (TASKS[0], TASKS[3]):
        n, T = 256, 1000
        env = b2sim.BatchedTaskEnv(env_id, n, seed=7, max_episode_steps=60000)
        _, model = oracle.load_urdf(model_files[model_name])
        ref = oracle.sample_reset_batch(task, 7, 0, n, 0)
        rng = np.random.default_rng(3)
        small = 2.0 if task == 1 else 5.0  # gentle actions so that no env terminates
        actions = make_actions(rng, T, n, small)
        elapsed = np.zeros(n, np.int32)
        _, _, d_ref = oracle.rollout(model, task, actions, ref, elapsed, max_episode_steps=60000, seed=7)
        a_dev = torch.as_tensor(actions, device="cuda")
        for t in range(T):
            env.step(a_dev[t])
        torch.cuda.synchronize()
        alive = d_ref.sum(axis=0) == 0
        assert alive.sum() > n // 4
        np.testing.assert_allclose(env.state.cpu().numpy()[alive], ref[alive], rtol=1e-6, atol=1e-9)
        env.close()


def test_action_repeat_steps_per_run(torch, oracle, model_files):
    """agent_rate < physics_rate: steps_per_run physics iterations per env step, the one-shot force acting on
    the first iteration only (Physics.cpp:2250-2254, tests/.python/test_joint_force.py:9-81)."""
    import b2sim
    n, T = 512, 100
    env = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=6, physics_rate=1000.0, agent_rate=250.0)
    assert env.sim.steps_per_run == 4
    _, model = oracle.load_urdf(model_files["cartpole"])
    ref = oracle.sample_reset_batch(4, 6, 0, n, 0)
    actions = make_actions(np.random.default_rng(8), T, n, 200.0)
    elapsed = np.zeros(n, np.int32)
    o_ref, r_ref, d_ref = oracle.rollout(model, 4, actions, ref, elapsed, seed=6, steps_per_run=4)
    a_dev = torch.as_tensor(actions, device="cuda")
    for t in range(T):
        o, r, d = env.step(a_dev[t])
        assert np.array_equal(d.cpu().numpy(), d_ref[t])
        np.testing.assert_allclose(o.cpu().numpy(), o_ref[t], rtol=1e-9, atol=1e-11)
    assert env.sim.time() == pytest.approx(T * 0.004)
    env.close()


def test_rollout_api_equals_step_loop(torch):
    import b2sim
    n, T = 4096, 64
    a = b2sim.BatchedTaskEnv("Pendulum-Gazebo-v0", n, seed=1)
    b = b2sim.BatchedTaskEnv("Pendulum-Gazebo-v0", n, seed=1)
    acts = (torch.rand((T, n), dtype=torch.float64, device="cuda") * 2 - 1) * 50
    for t in range(T):
        a.step(acts[t])
    b.rollout(acts)
    torch.cuda.synchronize()
    assert torch.equal(a.state, b.state) and torch.equal(a.obs, b.obs) and torch.equal(a.done, b.done)
    # and inside a CUDA graph: the Philox step index lives in a device counter that the kernel advances itself, so
    # replays keep drawing fresh reset states exactly like eager stepping does
    c = b2sim.BatchedTaskEnv("Pendulum-Gazebo-v0", n, seed=1)
    stream = torch.cuda.Stream()
    c.use_stream(stream)
    graph = torch.cuda.CUDAGraph()
    half = acts[: T // 2].clone()
    with torch.cuda.stream(stream):
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            c.rollout(half)
    torch.cuda.synchronize()
    graph.replay()                       # steps 1 .. T/2
    half.copy_(acts[T // 2:])
    graph.replay()                       # steps T/2+1 .. T with the second half of the actions
    torch.cuda.synchronize()
    assert (a.elapsed.int() < T).sum().item() > 0                           # resets happened along the way
    assert torch.equal(c.state, a.state) and torch.equal(c.elapsed, a.elapsed) and torch.equal(c.obs, a.obs)
    assert c.sim.lib.b2sim_task_steps_done(c.sim.handle, c.model) == T
    for e in (a, b, c):
        e.close()


def test_sharding_is_index_exact(torch, oracle):
    """Two shards with env_offset reproduce the unsharded run bit for bit (Philox keyed by global index)."""
    import b2sim
    n, T = 512, 120
    rng = np.random.default_rng(9)
    actions = torch.as_tensor(make_actions(rng, T, n, 200.0), device="cuda")
    full = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=4, max_episode_steps=50)
    lo = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n // 2, seed=4, max_episode_steps=50)
    hi = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n // 2, seed=4, env_offset=n // 2,
                              max_episode_steps=50)
    for t in range(T):
        full.step(actions[t])
        lo.step(actions[t, : n // 2].contiguous())
        hi.step(actions[t, n // 2:].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(full.state[: n // 2], lo.state) and torch.equal(full.state[n // 2:], hi.state)
    assert torch.equal(full.done[: n // 2], lo.done) and torch.equal(full.done[n // 2:], hi.done)
    for e in (full, lo, hi):
        e.close()


@pytest.mark.parametrize("n", [2048, 3 * 65536 + 17])
def test_step_host_matches_device_path(n, torch):
    """Host-buffer entry point (chunked: H2D, kernel and D2H of different env ranges overlap) against the
    single-launch device path, including envs that terminate and reset."""
    import b2sim
    a = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=2, max_episode_steps=12)
    b = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=2, max_episode_steps=12)
    rng = np.random.default_rng(1)
    obs = np.zeros((n, 4)); rew = np.zeros(n); done = np.zeros(n, np.uint8)
    for t in range(20):
        act = rng.uniform(-200, 200, n)
        a.step_host(act, obs, rew, done)
        o, r, d = b.step(torch.as_tensor(act, device="cuda"))
        torch.cuda.synchronize()
        assert np.array_equal(obs, o.cpu().numpy()) and np.array_equal(rew, r.cpu().numpy())
        assert np.array_equal(done, d.cpu().numpy())
    assert torch.equal(a.state, b.state) and torch.equal(a.elapsed, b.elapsed)
    a.close(); b.close()


def test_fp32_fast_mode_tracks_fp64(torch):
    """fp32 fast mode is reported separately; here it only has to stay close to fp64 over a short horizon."""
    import b2sim
    n, T = 1024, 50
    e64 = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=5)
    e32 = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=5, dtype="float32")
    np.testing.assert_allclose(e32.state.cpu().numpy(), e64.state.cpu().numpy(), rtol=1e-6, atol=1e-7)
    rng = np.random.default_rng(1)
    for t in range(T):
        act = rng.uniform(-50, 50, n).astype(np.float32)
        e64.step(torch.as_tensor(act.astype(np.float64), device="cuda"))
        e32.step(torch.as_tensor(act, device="cuda"))
    torch.cuda.synchronize()
    np.testing.assert_allclose(e32.state.cpu().numpy(), e64.state.cpu().numpy(), rtol=2e-3, atol=2e-3)
    e64.close(); e32.close()


@pytest.mark.parametrize("env_id,task,model_name,amp", TASKS)
def test_fp32_fast_mode_one_step_against_the_fp64_oracle(env_id, task, model_name, amp, torch, oracle, model_files):
    """The fp32 fast mode against the fp64 oracle, one step at a time from identical states (the device's fp32 state
    converted exactly to fp64): observations and rewards agree to a few 1e-5 (single-precision rounding through the
    sine / cosine and the 2x2 solve), done masks agree except for envs within 1e-4 of a termination bound. Forty
    consecutive steps, each re-anchored on the device state, so the bound is per step and not a drift allowance."""
    import b2sim
    n, seed = 2048, 9
    env = b2sim.BatchedTaskEnv(env_id, n, seed=seed, dtype="float32", max_episode_steps=5000)
    _, model = oracle.load_urdf(model_files[model_name])
    rng = np.random.default_rng(2)
    mismatched = 0
    for t in range(40):
        state = env.state.cpu().numpy().astype(np.float64)
        act = make_actions(rng, 1, n, amp)
        ref = state.copy()
        elapsed = env.elapsed.cpu().numpy().astype(np.int32)
        o_ref, r_ref, d_ref = oracle.rollout(model, task, act, ref, elapsed, max_episode_steps=5000, seed=seed, first_step=t + 1)
        obs, rew, done = env.step(torch.as_tensor(act[0].astype(np.float32), device="cuda"))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        same = done == d_ref[0]
        mismatched += int((~same).sum())
        np.testing.assert_allclose(obs[same], o_ref[0][same], rtol=3e-5, atol=3e-5, err_msg=f"step {t}")
        np.testing.assert_allclose(rew[same], r_ref[0][same], rtol=3e-5, atol=3e-5, err_msg=f"step {t}")
    assert mismatched <= 4, f"{mismatched} done flags differ between fp32 and fp64 (only envs on a bound may)"
    env.close()


# --------------------------------------------------------------------------------------------------
# generic tree kernel (GazeboSimulator::run) against the oracle's single-world simulator
# --------------------------------------------------------------------------------------------------
F_POS, F_VEL, F_ACC, F_FORCE, F_FT, F_PT, F_VT, F_PR, F_VR = range(9)
MODE_IDLE, MODE_FORCE, MODE_VELOCITY, MODE_FOLLOWER, MODE_POSITION = 1, 2, 3, 4, 5

PANDA_GAINS = [(50, 0, 20), (10000, 0, 500), (100, 0, 10), (1000, 0, 50), (100, 0, 10), (100, 0, 10), (10, 0.5, 0.1),
               (100, 0, 50), (100, 0, 50)]
PANDA_Q0 = [0, -0.785, 0, -2.356, 0, 1.571, 0.785, 0.0, 0.0]
DBL_MAX = np.finfo(np.float64).max


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "panda"])
def test_run_force_mode_matches_oracle(name, torch, oracle, model_files):
    import b2sim
    n, T = 8, 200
    sim = b2sim.Simulator(n, 0.001, 1)
    mid = sim.insert_model_file(model_files[name])
    _, model = oracle.load_urdf(model_files[name])
    nq = model.nb
    rng = np.random.default_rng(0)
    refs = [oracle.Sim(model, 0.001, 1) for _ in range(n)]
    for j in range(nq):
        sim.set_control_mode(mid, j, MODE_FORCE)
        for r in refs:
            r.set_control_mode(j, MODE_FORCE)
    q0 = rng.uniform(-0.5, 0.5, (n, nq)) + (np.array(PANDA_Q0) if name == "panda" else 0)
    dq0 = rng.uniform(-0.5, 0.5, (n, nq))
    for e in range(n):
        for j in range(nq):
            sim.set_joint(mid, F_PR, e, j, q0[e, j]); sim.set_joint(mid, F_VR, e, j, dq0[e, j])
            refs[e].reset_position(j, q0[e, j]); refs[e].reset_velocity(j, dq0[e, j])
    # resets are deferred to the next run (tests/test_scenario/test_model.py:82-94)
    assert sim.get_joint(mid, F_POS, 0, 0) == 0.0
    sim.run(paused=True)
    for r in refs:
        r.run(True)
    assert sim.time() == 0.0
    assert sim.get_joint(mid, F_POS, 3, 0) == q0[3, 0]
    state = sim.tensor(mid, 0)
    fcmd = sim.tensor(mid, 2)
    for t in range(T):
        tau = rng.uniform(-3, 3, (n, nq))
        fcmd.copy_(torch.as_tensor(tau, device="cuda"))
        sim.run()
        for e in range(n):
            for j in range(nq):
                refs[e].set_force_target(j, tau[e, j])
            refs[e].run(False)
    got = state.cpu().numpy()
    ref = np.array([[r.position(j) for j in range(nq)] + [r.velocity(j) for j in range(nq)] for r in refs])
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-10)
    assert sim.time() == pytest.approx(T * 0.001, abs=1e-12) and refs[0].time() == sim.time()
    # one-shot command: zero after the step (Physics.cpp:2250-2254)
    assert sim.get_joint(mid, F_FT, 0, 0) == 0.0
    sim.close()


def test_panda_position_pid_matches_oracle_and_holds_pose(torch, oracle, model_files):
    """Panda position PID @ 1 kHz, controller period = dt (tests/test_scenario/test_pid_controllers.py)."""
    import b2sim
    n, T = 4, 500
    sim = b2sim.Simulator(n, 0.001, 1)
    mid = sim.insert_model_file(model_files["panda"])
    _, model = oracle.load_urdf(model_files["panda"])
    ref = oracle.Sim(model, 0.001, 1)
    for j in range(9):
        sim.set_joint(mid, F_PR, -1, j, PANDA_Q0[j])
        ref.reset_position(j, PANDA_Q0[j])
    sim.run(paused=True); ref.run(True)
    sim.set_controller_period(mid, 0.001); ref.set_controller_period(0.001)
    for j, (p, i, d) in enumerate(PANDA_GAINS):
        sim.set_pid(mid, j, p, i, d, DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX, 0.0)
        ref.set_pid(j, p, i, d, DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX, 0.0)
        sim.set_control_mode(mid, j, MODE_POSITION); ref.set_control_mode(j, MODE_POSITION)
    # target seeded with the current position (Joint.cpp:430-435)
    assert sim.get_joint(mid, F_PT, 1, 3) == PANDA_Q0[3]
    for t in range(T):
        target = PANDA_Q0[0] + 0.3 * np.sin(2 * np.pi * 0.33 * t * 0.001)
        sim.set_joint(mid, F_PT, -1, 0, target); ref.set_position_target(0, target)
        sim.run(); ref.run(False)
    got = sim.tensor(mid, 0).cpu().numpy()
    want = np.array([ref.position(j) for j in range(9)] + [ref.velocity(j) for j in range(9)])
    for e in range(n):
        np.testing.assert_allclose(got[e], want, rtol=1e-7, atol=1e-9)
    # holds the pose within 1 degree on the joints that are not driven
    assert np.all(np.abs(got[0, 1:7] - np.array(PANDA_Q0[1:7])) < np.deg2rad(3.0))  # test_pid_controllers.py:112-115
    sim.close()


def test_kindyn_matches_oracle(torch, oracle, model_files):
    import b2sim
    n = 64
    sim = b2sim.Simulator(n, 0.001, 1)
    mid = sim.insert_model_file(model_files["panda"], pose=(0.1, -0.2, 0.3, 1, 0, 0, 0))
    t, model = oracle.load_urdf(model_files["panda"], base_position=(0.1, -0.2, 0.3))
    D = oracle.Dynamics(model)
    rng = np.random.default_rng(4)
    q = rng.uniform(-1, 1, (n, 9)) + np.array(PANDA_Q0)
    dq = rng.uniform(-1, 1, (n, 9))
    sim.tensor(mid, 0).copy_(torch.as_tensor(np.concatenate([q, dq], 1), device="cuda"))
    info = sim.info(mid)
    ee = info.link_names.index("end_effector_frame")
    M = torch.empty((n, 81), dtype=torch.float64, device="cuda")
    h = torch.empty((n, 9), dtype=torch.float64, device="cuda")
    J = torch.empty((n, 54), dtype=torch.float64, device="cuda")
    sim.kindyn(mid, ee, M, h, J)
    sim.update_kinematics(mid)
    poses = sim.tensor(mid, 13).cpu().numpy().reshape(n, -1, 7)
    torch.cuda.synchronize()
    body = int(t["link_body"][t["link_names"].index("end_effector_frame")])
    off_p = t["link_p"][t["link_names"].index("end_effector_frame")]
    for e in range(0, n, 7):
        np.testing.assert_allclose(M[e].cpu().numpy().reshape(9, 9), D.mass_matrix(q[e]), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(h[e].cpu().numpy(), D.inverse_dynamics(q[e], dq[e], np.zeros(9)), rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(J[e].cpu().numpy().reshape(6, 9), D.point_jacobian(q[e], body, off_p), rtol=1e-10, atol=1e-12)
        Rw, pw = D.forward_kinematics(q[e])
        np.testing.assert_allclose(poses[e, ee, :3], pw[body] + Rw[body] @ off_p, rtol=1e-10, atol=1e-12)
    sim.close()


def test_chain_closed_form_equals_tree_kernel(torch, model_files):
    """The fused closed-form cartpole step and the generic articulated-body kernel agree to rounding."""
    import b2sim
    n = 1024
    env = b2sim.BatchedTaskEnv("CartPoleContinuousSwingup-Gazebo-v0", n, seed=3)
    sim = b2sim.Simulator(n, 0.001, 1)
    mid = sim.insert_model_file(model_files["cartpole"])
    sim.set_control_mode(mid, 0, MODE_FORCE)
    sim.tensor(mid, 0).copy_(env.state)
    rng = np.random.default_rng(0)
    for t in range(30):
        act = torch.as_tensor(rng.uniform(-100, 100, n), device="cuda")
        sim.tensor(mid, 2)[:, 0].copy_(act)
        sim.run()
        env.step(act)
    torch.cuda.synchronize()
    alive = env.elapsed.cpu().numpy() == 30
    assert alive.sum() > n // 2
    np.testing.assert_allclose(sim.tensor(mid, 0).cpu().numpy()[alive], env.state.cpu().numpy()[alive], rtol=1e-9, atol=1e-12)
    env.close(); sim.close()


def test_panda_fused_task_matches_oracle(torch, oracle, model_files):
    """BASELINE config 4: fused Panda kernel (position PID + articulated-body step with the fingers resting on
    their joint limits + end-effector pose / Jacobian observation) against the oracle's single-world simulator
    and its FK / Jacobian, env by env."""
    import b2sim
    from b2sim.batched import PANDA_PID, PANDA_Q0
    n, T = 8, 120
    env = b2sim.BatchedTaskEnv("PandaReach-Gazebo-v0", n, max_episode_steps=100)
    assert env.nobs == 115 and env.nact == 9
    t, model = oracle.load_urdf(model_files["panda"])
    D = oracle.Dynamics(model)
    l = t["link_names"].index("end_effector_frame")
    body, off_p = int(t["link_body"][l]), t["link_p"][l]

    def fresh_ref():
        r = oracle.Sim(model, 0.001, 1)
        for j in range(9):
            r.reset_position(j, PANDA_Q0[j])
        r.run(True)
        r.set_controller_period(0.001)
        for j, (p, i, d) in enumerate(PANDA_PID):
            r.set_pid(j, p, i, d, DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX, 0.0)
            r.set_control_mode(j, MODE_POSITION)
        return r

    refs = [fresh_ref() for _ in range(n)]
    rng = np.random.default_rng(3)
    phase = rng.uniform(0, 6.28, (n, 1))
    for step in range(T):
        wave = np.sin(2 * np.pi * 0.33 * step * 0.001 + phase)
        targets = np.array(PANDA_Q0) + 0.1 * wave * np.ones((n, 9))
        # fingers: open to mid-range and stay inside (0, 0.04). A joint sitting exactly on a limit switches its
        # limit row on q <= lower, a knife edge at rounding-noise level (DART's own activation rule), so exact
        # trajectory parity is only meaningful away from it; holding at the limit is checked further down.
        targets[:, 7:] = 0.02 + 0.01 * wave
        obs, rew, done = env.step(torch.as_tensor(targets, device="cuda"))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        for e in range(n):
            r = refs[e]
            for j in range(9):
                r.set_position_target(j, targets[e, j])
            r.run(False)
            q = np.array([r.position(j) for j in range(9)]); dq = np.array([r.velocity(j) for j in range(9)])
            np.testing.assert_allclose(obs[e, :9], q, rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(obs[e, 9:18], dq, rtol=1e-7, atol=1e-9)
            Rw, pw = D.forward_kinematics(q)
            pe = pw[body] + Rw[body] @ off_p
            np.testing.assert_allclose(obs[e, 18:21], pe, rtol=1e-8, atol=1e-10)
            J = obs[e, 25:].reshape(6, 15)
            np.testing.assert_allclose(J[:, 6:], D.point_jacobian(q, body, off_p), rtol=1e-7, atol=1e-9)
            np.testing.assert_allclose(J[:3, :3], np.eye(3)); np.testing.assert_allclose(J[3:, 3:6], np.eye(3))
            assert rew[e] == pytest.approx(-np.linalg.norm(pe - np.array([0.5, 0.0, 0.5])), rel=1e-8)
        if step == 99:   # TimeLimit -> reset to the initial configuration, PID state cleared
            assert done.sum() == n
            np.testing.assert_allclose(env.state.cpu().numpy(), np.tile(PANDA_Q0 + [0.0] * 9, (n, 1)))
            assert int(env.elapsed.cpu().numpy().max()) == 0
            refs = [fresh_ref() for _ in range(n)]
        else:
            assert done.sum() == 0
    # fingers driven against their lower limit: the limit rows keep them there (velocity-level constraint:
    # at most one step of travel can slip through when a row toggles)
    env.reset()
    for step in range(300):
        targets = np.tile(PANDA_Q0, (n, 1))
        targets[:, 7:] = -0.05
        obs, _, _ = env.step(torch.as_tensor(targets, device="cuda"))
    fingers = obs[:, 7:9].cpu().numpy()
    assert fingers.min() > -1e-3 and fingers.max() < 1e-3
    env.close()


def test_tree_kernel_constraint_paths_match_oracle(torch, oracle, model_files):
    """Joint limits through the articulated-body impulse path (cart pushed into the end of the rail, no damping)
    and through the dense M^-1 path (damped model), step by step against the oracle."""
    import b2sim
    for damping in ("0.0", "0.3"):
        xml = open(model_files["cartpole"]).read().replace('damping="0.0"', f'damping="{damping}"')
        _, model = oracle.load_urdf(xml)
        n, T = 4, 2500
        sim = b2sim.Simulator(n, 0.001, 1)
        mid = sim.insert_model(xml)
        sim.set_control_mode(mid, 0, MODE_FORCE)
        refs = [oracle.Sim(model, 0.001, 1) for _ in range(n)]
        for r in refs:
            r.set_control_mode(0, MODE_FORCE)
        fcmd, state = sim.tensor(mid, 2), sim.tensor(mid, 0)
        force = np.array([40.0, -40.0, 25.0, 60.0])
        for step in range(T):
            fcmd[:, 0].copy_(torch.as_tensor(force, device="cuda"))
            sim.run()
            for e, r in enumerate(refs):
                r.set_force_target(0, force[e])
                r.run(False)
        got = state.cpu().numpy()
        ref = np.array([[r.position(0), r.position(1), r.velocity(0), r.velocity(1)] for r in refs])
        assert np.all(np.abs(ref[:, 0]) >= 2.6)            # every cart reached the end of the rail
        np.testing.assert_allclose(got[:, 0], ref[:, 0], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(got[:, 2], ref[:, 2], rtol=0, atol=1e-8)
        sim.close()


@pytest.mark.parametrize("env_id,task,model_name,amp", [TASKS[0], TASKS[3]])
def test_domain_randomization_matches_oracle(env_id, task, model_name, amp, torch, oracle, model_files):
    """Per-env link masses and gravity (randomizers/cartpole.py:51-56,100-135): the kernel rebuilds the closed-form
    coefficients from (mass offsets, gravity scale); the oracle runs the articulated-body algorithm on a model
    with those masses and that gravity. Parameters are redrawn at every reset."""
    import b2sim
    n, T, seed = 512, 300, 21
    env = b2sim.BatchedTaskEnv(env_id, n, seed=seed, max_episode_steps=80)
    rand = env.randomize(mass_delta=0.2, gravity_sigma=0.2)
    nq = oracle.task_nq(task)
    assert rand.shape == (n, nq + 1) and env.bytes_per_env_step == (74 if task == 1 else 114) + 8 * (nq + 1)
    _, model = oracle.load_urdf(model_files[model_name])
    ref_rand = np.array([oracle.sample_rand_params(model, seed, e, 0, 0.2, 0.2) for e in range(n)])
    np.testing.assert_allclose(rand.cpu().numpy(), ref_rand, rtol=1e-12, atol=1e-14)
    assert 0.15 < np.abs(ref_rand[:, 0]).max() <= 0.2 and 0.005 < ref_rand[:, nq].std() < 0.05
    ref_state = oracle.sample_reset_batch(task, seed, 0, n, 0)
    assert np.array_equal(env.state.cpu().numpy(), ref_state)
    actions = make_actions(np.random.default_rng(4), T, n, amp)
    elapsed = np.zeros(n, np.int32)
    o_ref, r_ref, d_ref = oracle.rollout_randomized(model, task, actions, ref_state, elapsed, ref_rand, 0.2, 0.2,
                                                    max_episode_steps=80, seed=seed)
    a_dev = torch.as_tensor(actions, device="cuda")
    for t in range(T):
        o, r, d = env.step(a_dev[t])
        assert np.array_equal(d.cpu().numpy(), d_ref[t]), t
        np.testing.assert_allclose(o.cpu().numpy(), o_ref[t], rtol=1e-9, atol=1e-11)
    assert d_ref.sum() > n
    np.testing.assert_allclose(env.state.cpu().numpy(), ref_state, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(rand.cpu().numpy(), ref_rand, rtol=1e-12, atol=1e-14)
    # and it changes the dynamics: the same rollout without randomisation ends elsewhere
    plain = b2sim.BatchedTaskEnv(env_id, n, seed=seed, max_episode_steps=80)
    for t in range(T):
        plain.step(a_dev[t])
    assert (plain.state - env.state).abs().max().item() > 1e-3
    env.close(); plain.close()


def test_fp32_fast_mode_tree_panda_and_contacts(torch, model_files):
    """fp32 fast mode of the tree, Panda and contact kernels stays close to fp64 over a short horizon (it is
    reported separately and carries no 1e-9 claim)."""
    import b2sim
    from b2sim.batched import PANDA_Q0
    n = 256
    envs = {dt: b2sim.BatchedTaskEnv("PandaReach-Gazebo-v0", n, dtype=dt) for dt in ("float64", "float32")}
    rng = np.random.default_rng(0)
    phase = rng.uniform(0, 6.28, (n, 1))
    for step in range(100):
        wave = np.sin(2 * np.pi * 0.33 * step * 0.001 + phase)
        targets = np.array(PANDA_Q0) + 0.1 * wave * np.ones((n, 9))
        targets[:, 7:] = 0.02 + 0.01 * wave
        envs["float64"].step(torch.as_tensor(targets, device="cuda"))
        envs["float32"].step(torch.as_tensor(targets.astype(np.float32), device="cuda"))
    o64, o32 = envs["float64"].obs.cpu().numpy(), envs["float32"].obs.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(o32[:, :9], o64[:, :9], atol=2e-3)       # joint positions
    np.testing.assert_allclose(o32[:, 18:21], o64[:, 18:21], atol=2e-3)  # end-effector position
    for e in envs.values():
        e.close()
    # free bodies + contacts
    cube = open(model_files["ground_plane"]).read()
    from test_contacts_gpu import CUBE_URDF
    finals = {}
    for dt in ("float64", "float32"):
        sim = b2sim.Simulator(64, 0.001, 1, dt)
        sim.insert_model(cube)
        c = sim.insert_model(CUBE_URDF, pose=(0, 0, 0.15, 1, 0, 0, 0), name="cube")
        for _ in range(300):
            sim.run()
        finals[dt] = sim.tensor(c, 14).cpu().numpy().astype(np.float64)
        assert len(sim.contacts(0)) == 4
        sim.close()
    np.testing.assert_allclose(finals["float32"][:, 2], finals["float64"][:, 2], atol=2e-3)
    assert abs(finals["float32"][:, 2].mean() - 0.1) < 3e-3


@pytest.mark.parametrize("env_id,task,model_name,amp", TASKS)
@pytest.mark.parametrize("randomize", [False, True])
def test_single_launch_trajectory_equals_the_step_loop(env_id, task, model_name, amp, randomize, torch, oracle, model_files):
    """b2sim_task_trajectory (k_task_trajectory: T env.steps in one launch, state in registers) against T launches of
    k_task_chain on the same seeds and actions: done masks, episode counters, reset states and redrawn randomised
    parameters bit for bit; observations / rewards / states to rounding (1e-9: the two kernels inline the same step,
    but the compiler contracts its FMAs per kernel); then stepping continues from it. The recorded trajectory is
    also checked against the CPU oracle (1e-9)."""
    import b2sim
    n, T, seed, offset = 1000, 333, 21, 77  # n not a multiple of the block, T not a multiple of the prefetch depth
    envs = [b2sim.BatchedTaskEnv(env_id, n, seed=seed, env_offset=offset, max_episode_steps=120) for _ in range(2)]
    rands = [e.randomize(0.2, 0.2) for e in envs] if randomize else None
    rng = np.random.default_rng(9)
    actions = make_actions(rng, T + 5, n, amp)
    a_dev = torch.as_tensor(actions, device="cuda")
    loop, fused = envs
    obs = torch.empty((T, n, loop.nobs), dtype=torch.float64, device="cuda")
    rew = torch.empty((T, n), dtype=torch.float64, device="cuda")
    done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    for t in range(T):
        o, r, d = loop.step(a_dev[t])
        obs[t].copy_(o); rew[t].copy_(r); done[t].copy_(d)
    o2, r2, d2 = fused.trajectory(a_dev[:T].contiguous())
    torch.cuda.synchronize()

    def same(tag):
        assert torch.equal(fused.done, loop.done) and torch.equal(fused.elapsed, loop.elapsed), tag
        for name in ("state", "obs", "reward"):
            torch.testing.assert_close(getattr(fused, name), getattr(loop, name), rtol=1e-9, atol=1e-11, msg=f"{tag}: {name}")
        if randomize:  # redrawn at the same resets from the same Philox stream
            assert torch.equal(rands[0], rands[1]), tag

    assert done.sum().item() > 0
    assert torch.equal(d2, done)
    torch.testing.assert_close(o2, obs, rtol=1e-9, atol=1e-11)
    torch.testing.assert_close(r2, rew, rtol=1e-9, atol=1e-11)
    same("after the rollout")
    # both continue identically: step indices and (randomised) parameters are in the same place
    fused.trajectory(a_dev[T:T + 3].contiguous(), record=False)
    for t in range(T, T + 3):
        loop.step(a_dev[t])
    for t in range(T + 3, T + 5):
        loop.step(a_dev[t]); fused.step(a_dev[t])
    torch.cuda.synchronize()
    same("after continuing")
    if not randomize:
        _, model = oracle.load_urdf(model_files[model_name])
        ref_state = oracle.sample_reset_batch(task, seed, offset, n, 0)
        elapsed = np.zeros(n, np.int32)
        o_ref, r_ref, d_ref = oracle.rollout(model, task, actions[:T], ref_state, elapsed, max_episode_steps=120, seed=seed,
                                             env_offset=offset, first_step=1)
        assert np.array_equal(d2.cpu().numpy(), d_ref)
        np.testing.assert_allclose(o2.cpu().numpy(), o_ref, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(r2.cpu().numpy(), r_ref, rtol=1e-9, atol=1e-11)
    for e in envs:
        e.close()
