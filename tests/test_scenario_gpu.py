"""The drop-in ``scenario`` module and the runtimes on the B200: re-statements of the reference's own tests
(tests/test_scenario/*.py, tests/test_gym_ignition/*.py) plus the single-env gym flow checked against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def default_world():
    """tests/common/utils.py:53-82 default_world_fixture: initialized simulator + ground plane + DART."""
    from gym_ignition.utils.scenario import init_gazebo_sim
    gazebo, world = init_gazebo_sim(step_size=0.001, real_time_factor=1.0, steps_per_run=1)
    yield gazebo, world
    gazebo.close()


def test_simulator_configuration_is_validated():
    """tests/test_scenario/test_gazebo_simulator.py:17-51."""
    from scenario import gazebo as scenario
    for args in ((0.0, 1.0, 1), (-0.001, 1.0, 1), (0.001, 0.0, 1), (0.001, 1.0, 0)):
        sim = scenario.GazeboSimulator(*args)
        assert not sim.initialize() and not sim.initialized()
    sim = scenario.GazeboSimulator(0.001, 1.0, 1)
    assert not sim.run()
    assert sim.initialize() and sim.initialized()
    assert sim.world_names() == ("default",)
    assert sim.step_size() == 0.001 and sim.steps_per_run() == 1 and sim.real_time_factor() == 1.0
    assert sim.close()


def test_world_api(model_files):
    """tests/test_scenario/test_world.py:18-219."""
    from scenario import core
    from scenario import gazebo as scenario
    sim = scenario.GazeboSimulator(0.001, 1.0, 1)
    assert sim.initialize()
    world = sim.get_world()
    assert world.name() == "default" and world.id() != 0 and world.valid()
    assert world.time() == 0.0 and world.model_names() == ()
    assert world.gravity() == pytest.approx((0, 0, -9.8))
    assert world.insert_model(model_files["ground_plane"])
    assert not world.insert_model(model_files["ground_plane"])            # name clash
    assert world.insert_model(model_files["ground_plane"], core.Pose_identity(), "other_plane")
    assert not world.insert_model("/does/not/exist.urdf")
    pose = core.Pose([0.5, -1.0, 0.25], [1.0, 0, 0, 0])
    assert world.insert_model_from_string(open(model_files["pendulum"]).read(), pose, "pend")
    assert set(world.model_names()) == {"ground_plane", "other_plane", "pend"}
    with pytest.raises(RuntimeError):
        world.get_model("missing")
    model = world.get_model("pend")
    assert model is world.get_model("pend")                                # cached objects
    assert model.base_position() == pytest.approx((0.5, -1.0, 1.25))      # support link sits 1 m above the pose
    # without physics the world time never advances (test_world.py:120-146)
    assert sim.run() and world.time() == 0.0
    # removal is deferred to the next run
    assert world.remove_model("other_plane")
    assert "other_plane" in world.model_names()
    assert sim.run(paused=True)
    assert "other_plane" not in world.model_names()
    assert not world.remove_model("other_plane")
    # once physics is loaded, the time catches up with the server (test_world.py:177-188)
    assert world.set_physics_engine(scenario.PhysicsEngine_dart)
    assert sim.run(paused=True) and world.time() == pytest.approx(0.001)
    for k in range(3):
        assert sim.run()
    assert world.time() == pytest.approx(0.004)
    assert sim.run(paused=True) and world.time() == pytest.approx(0.004)
    assert not world.set_gravity((0, 0, -5.0))                             # World.cpp:301-319
    sim.close()


def test_multi_world(model_files):
    """tests/test_scenario/test_multi_world.py: several worlds per simulator, unique names and ids."""
    from scenario import gazebo as scenario
    sim = scenario.GazeboSimulator(0.001, 1.0, 1)
    assert sim.insert_world_from_sdf("", "a") and sim.insert_world_from_sdf("", "b")
    assert not sim.insert_world_from_sdf("", "a")
    assert sim.initialize()
    assert set(sim.world_names()) == {"a", "b"}
    wa, wb = sim.get_world("a"), sim.get_world("b")
    assert wa.id() != wb.id()
    with pytest.raises(RuntimeError):
        sim.get_world()
    assert set(scenario.ECMSingleton_instance().world_names()) >= {"a", "b"}
    for w in (wa, wb):
        assert w.insert_model(model_files["pendulum"]) and w.set_physics_engine(scenario.PhysicsEngine_dart)
    wa.get_model("pendulum").get_joint("pivot").reset_position(0.1)
    for _ in range(10):
        sim.run()
    assert wa.get_model("pendulum").get_joint("pivot").position() != 0.0
    assert wb.get_model("pendulum").get_joint("pivot").position() == 0.0   # upright equilibrium, untouched
    sim.close()
    assert "a" not in scenario.ECMSingleton_instance().world_names()


def test_model_joint_api(default_world, model_files):
    """tests/test_scenario/test_model.py:69-110,221-302."""
    from scenario import core
    gazebo, world = default_world
    assert world.insert_model(model_files["panda"])
    panda = world.get_model("panda").to_gazebo()
    assert panda.dofs() == 9 and panda.nr_of_joints() == 9 and panda.name() == "panda"
    assert panda.total_mass() == pytest.approx(sum(l.mass() for l in panda.links()))
    assert panda.joint_positions() == pytest.approx([0.0] * 9)
    with pytest.raises(RuntimeError):
        panda.get_joint("nope")
    with pytest.raises(RuntimeError):
        panda.joint_position_targets()                                     # never set: ComponentNotFound
    # resets are deferred to the next run
    q = [0.1 * k for k in range(9)]
    assert panda.reset_joint_positions(q)
    assert panda.joint_positions() == pytest.approx([0.0] * 9)
    assert gazebo.run(paused=True)
    assert panda.joint_positions() == pytest.approx(q)
    # subset with custom serialization
    names = ["panda_joint6", "panda_joint2"]
    assert panda.reset_joint_velocities([0.5, -0.5], names)
    assert gazebo.run(paused=True)
    assert panda.joint_velocities(names) == pytest.approx([0.5, -0.5])
    assert panda.joint_positions(names) == pytest.approx([q[5], q[1]])
    assert not panda.reset_joint_positions([0.0], names)                  # size mismatch
    # target setters / getters in Force mode
    assert panda.set_joint_control_mode(core.JointControlMode_force)
    assert panda.joint_generalized_force_targets() == pytest.approx([0.0] * 9)
    assert panda.set_joint_position_targets([0.2] * 9) and panda.joint_position_targets() == pytest.approx([0.2] * 9)
    assert panda.set_joint_velocity_targets([0.3] * 9) and panda.joint_velocity_targets() == pytest.approx([0.3] * 9)
    assert panda.set_joint_generalized_force_targets([1.0] * 9)
    assert panda.joint_generalized_force_targets() == pytest.approx([1.0] * 9)
    # history of applied forces: 3 steps x 9 DoF, time-major
    assert panda.enable_history_of_applied_joint_forces(True, 3)
    for k in range(1, 5):
        assert panda.set_joint_generalized_force_targets([float(k)] * 9)
        assert gazebo.run()
    hist = panda.history_of_applied_joint_forces()
    assert len(hist) == 27 and hist == pytest.approx([2.0] * 9 + [3.0] * 9 + [4.0] * 9)
    assert panda.joint_generalized_force_targets() == pytest.approx([0.0] * 9)   # one-shot commands
    j = panda.get_joint("panda_joint4")
    assert j.position_limit().min == pytest.approx(-3.0718) and j.max_generalized_force() == 87
    assert j.type() == core.JointType_revolute and panda.get_joint("panda_finger_joint1").type() == core.JointType_prismatic
    assert not j.set_max_generalized_force(100.0)                          # frozen after the first step


def test_position_pid_reference_test(default_world, model_files):
    """tests/test_scenario/test_pid_controllers.py:33-115 (shortened tracking phase)."""
    from scenario import core
    from gym_ignition_environments.models.panda import PID_GAINS_1000HZ
    gazebo, world = default_world
    assert world.insert_model(model_files["panda"])
    panda = world.get_model("panda").to_gazebo()
    joint1, joint6 = panda.get_joint("panda_joint1"), panda.get_joint("panda_joint6")
    r1 = abs(joint1.position_limit().max - joint1.position_limit().min)
    r6 = abs(joint6.position_limit().max - joint6.position_limit().min)
    assert joint1.reset_position(joint1.position_limit().min + r1 / 2)
    assert joint6.reset_position(joint6.position_limit().min + r6 / 2)
    assert gazebo.run(paused=True)
    assert panda.set_controller_period(gazebo.step_size())
    assert set(panda.joint_names()) == set(PID_GAINS_1000HZ)
    for name, (p, i, d) in PID_GAINS_1000HZ.items():
        assert panda.get_joint(name).set_pid(core.PID(p, i, d))
    # limits looser than the effort limit are replaced by +-effort (Joint.cpp:503-513)
    assert panda.get_joint("panda_joint1").pid().cmd_max == 87
    assert panda.set_joint_control_mode(core.JointControlMode_position)
    assert panda.joint_position_targets() == pytest.approx(panda.joint_positions())
    for _ in range(1000):
        assert gazebo.run()
    assert panda.joint_positions() == pytest.approx(panda.joint_position_targets(), abs=np.deg2rad(1))
    q01, q06 = joint1.position(), joint6.position()
    for k in range(1500):
        s = np.sin(2 * np.pi * 0.33 * k * 0.001)
        ref1, ref6 = q01 + 0.9 * r1 / 2 * s, q06 + 0.9 * r6 / 2 * s
        assert joint1.set_position_target(ref1) and joint6.set_position_target(ref6)
        assert gazebo.run()
        if k % 25 == 0:
            assert joint1.position() == pytest.approx(ref1, abs=np.deg2rad(3))
            assert joint6.position() == pytest.approx(ref6, abs=np.deg2rad(3))


def test_velocity_follower_and_friction(default_world, model_files):
    """tests/test_scenario/test_velocity_direct.py:19-82."""
    from scenario import core
    gazebo, world = default_world
    xml = open(model_files["pendulum"]).read().replace('damping="0.0" friction="0.0"', 'damping="0.2" friction="0.01"')
    assert world.insert_model_from_string(xml, core.Pose_identity(), "pendulum")
    pivot = world.get_model("pendulum").get_joint("pivot")
    assert pivot.coulomb_friction() == 0.01 and pivot.viscous_friction() == 0.2
    assert pivot.reset_position(np.deg2rad(90)) and gazebo.run(paused=True)
    for _ in range(5000):
        gazebo.run()
    assert abs(np.rad2deg(pivot.position())) % 360 == pytest.approx(180, abs=0.3)
    assert world.insert_model_from_string(xml, core.Pose([0, 1, 0], [1, 0, 0, 0]), "pendulum2")
    pivot2 = world.get_model("pendulum2").get_joint("pivot")
    assert pivot2.set_control_mode(core.JointControlMode_velocity_follower_dart)
    assert pivot2.set_velocity_target(np.pi) and gazebo.run()
    assert pivot2.velocity() == pytest.approx(np.pi, rel=1e-6)
    assert pivot2.set_velocity_target(-np.pi) and gazebo.run()
    assert pivot2.velocity() == pytest.approx(-np.pi, rel=1e-6)


def test_link_velocity_is_consistent_with_finite_differences(default_world, model_files):
    """tests/test_scenario/test_link_velocities.py:86-318 (velocity part): world linear velocity of a link
    equals the finite difference of its position; body = R^T world."""
    from scenario import core
    gazebo, world = default_world
    assert world.insert_model(model_files["panda"])
    panda = world.get_model("panda")
    assert panda.set_joint_control_mode(core.JointControlMode_force)
    rng = np.random.default_rng(0)
    assert panda.to_gazebo().reset_joint_velocities(list(rng.uniform(-0.5, 0.5, 9)))
    link = panda.get_link("panda_link7")
    gazebo.run(paused=True)
    for _ in range(20):
        p0 = np.array(link.position())
        gazebo.run()
        p1 = np.array(link.position())
        v = np.array(link.world_linear_velocity())
        assert (p1 - p0) / 0.001 == pytest.approx(v, abs=1e-2)
    R = np.array(_quat_R(link.orientation()))
    assert np.array(link.body_linear_velocity()) == pytest.approx(R.T @ np.array(link.world_linear_velocity()), abs=1e-12)


def _quat_R(q):
    w, x, y, z = q
    return [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]


@pytest.mark.parametrize("env_id,task", [("CartPoleDiscreteBalancing-Gazebo-v0", 2),
                                          ("CartPoleContinuousSwingup-Gazebo-v0", 4)])
def test_single_env_gym_flow_matches_oracle(env_id, task, oracle, model_files):
    """BASELINE config 1: the registered env through gym.make + the no-randomization wrapper
    (examples/python/launch_cartpole.py:32-75), single env, checked step by step against the oracle driven
    with the same actions and the same reset states."""
    import functools
    import gym
    import gym_ignition_environments  # noqa: F401  (registers the ids)
    from gym_ignition_environments import randomizers

    def make_env_from_id(env_id: str, **kwargs):
        return gym.make(env_id, **kwargs)

    env = randomizers.cartpole_no_rand.CartpoleEnvNoRandomizations(env=functools.partial(make_env_from_id, env_id=env_id))
    env.seed(42)
    _, model = oracle.load_urdf(model_files["cartpole"])
    for episode in range(2):
        obs = env.reset()
        ref = oracle.Sim(model, 0.001, 1)
        ref.set_control_mode(0, 2)
        x, dx, q, dq = obs
        ref.reset_position(0, x); ref.reset_position(1, q); ref.reset_velocity(0, dx); ref.reset_velocity(1, dq)
        ref.run(True)
        done, steps = False, 0
        while not done and steps < 300:
            action = env.action_space.sample()
            obs, reward, done, info = env.step(action)
            force, _ = oracle.action_force(task, float(np.asarray(action).ravel()[0]))
            ref.set_force_target(0, force)
            ref.run(False)
            st = [ref.position(0), ref.position(1), ref.velocity(0), ref.velocity(1)]
            o_ref, r_ref, d_ref = oracle.task_evaluate(task, st)
            np.testing.assert_allclose(obs, o_ref, rtol=1e-9, atol=1e-12)
            assert reward == pytest.approx(r_ref, rel=1e-9, abs=1e-12) and done == d_ref
            steps += 1
        assert steps > 5
        assert env.unwrapped.timestamp() > 0
    env.close()


def test_pendulum_runtime_reward_reads_zeroed_force_target(model_files):
    """Appendix A.2: after run() the force target reads 0, so the pendulum reward has no torque term."""
    import gym
    import gym_ignition_environments  # noqa: F401
    from gym_ignition_environments.models import pendulum
    env = gym.make("Pendulum-Gazebo-v0")
    runtime = env.unwrapped
    model = pendulum.Pendulum(world=runtime.world)
    runtime.task.model_name = model.name()
    runtime.gazebo.run(paused=True)
    env.seed(7)
    obs = env.reset()
    assert obs.shape == (3,) and abs(obs[0] ** 2 + obs[1] ** 2 - 1) < 1e-6
    obs, reward, done, _ = env.step(np.array([50.0], dtype=np.float32))
    q = model.get_joint("pivot").position(); dq = model.get_joint("pivot").velocity()
    assert model.get_joint("pivot").generalized_force_target() == 0.0
    assert reward == pytest.approx(-((100.0 if done else 0.0) + q * q + 0.1 * dq * dq), rel=1e-12)
    env.close()


@pytest.mark.parametrize("task", [1, 2, 3, 4])
def test_kernel_task_maths_match_reference_golden(task):
    """The fused kernels' observation / reward / done on the exact states of tests/golden (produced by the
    reference's Python task classes): done bit-exact, observations bit-exact for cartpole, rewards to 2 ulp."""
    import torch
    import b2sim
    G = np.load(os.path.join(HERE, "golden", "task_golden.npz"))
    states, obs, rew, done = (G[f"task{task}_{k}"] for k in ("states", "obs", "reward", "done"))
    env_id = {v[0]: k for k, v in b2sim.batched.TASKS.items()}[task]
    env = b2sim.BatchedTaskEnv(env_id, len(states))
    env.state.copy_(torch.as_tensor(states, device="cuda"))
    env.observe()
    torch.cuda.synchronize()
    got_o, got_r, got_d = env.obs.cpu().numpy(), env.reward.cpu().numpy(), env.done.cpu().numpy()
    assert np.array_equal(got_d, done)
    ok = ~np.isnan(states).any(axis=1)
    if task == 1:
        np.testing.assert_allclose(got_o[ok], obs[ok], rtol=0, atol=2.3e-16)
    else:
        assert np.array_equal(got_o[ok], obs[ok])
    scale = np.maximum(np.abs(rew[ok]), 1e-300)
    assert np.all(np.abs(got_r[ok] - rew[ok]) <= 2 * np.spacing(scale) + (2.3e-16 if task in (1, 4) else 0))
    env.close()


def test_batched_runtime_from_task_class():
    """BatchedGazeboRuntime drives the task's fused kernel and reproduces N single-env runtimes' semantics."""
    import torch
    from gym_ignition.runtimes.batched_runtime import BatchedGazeboRuntime
    from gym_ignition_environments.tasks.cartpole_continuous_swingup import CartPoleContinuousSwingup
    rt = BatchedGazeboRuntime(CartPoleContinuousSwingup, num_envs=4096, seed=3)
    assert rt.action_space.shape == (1,) and rt.observation_space.shape == (4,)
    rt.reset()
    for k in range(10):
        obs, rew, done = rt.step(torch.zeros(4096, dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    assert obs.shape == (4096, 4) and rew.shape == (4096,) and done.dtype == torch.uint8
    assert rt.timestamp() == pytest.approx(0.010)
    assert torch.all(torch.abs(obs[:, 2] - np.pi) < np.deg2rad(61) + 0.05)   # still hanging near the bottom
    rt.close()


def test_kindyn_wrapper(model_files, oracle):
    """KinDynComputations-style queries (rbd/idyntree/kindyncomputations.py:169-196,270-303,367-377)."""
    from gym_ignition.rbd.kindyn import KinDynComputations
    kd = KinDynComputations(model_files["panda"], world_gravity=np.array([0, 0, -9.806]))
    t, model = oracle.load_urdf(model_files["panda"], gravity=(0, 0, -9.806))
    D = oracle.Dynamics(model)
    rng = np.random.default_rng(0)
    s, ds = rng.uniform(-1, 1, 9), rng.uniform(-1, 1, 9)
    kd.set_robot_state(s, ds)
    np.testing.assert_allclose(kd.get_joint_positions(), s)
    np.testing.assert_allclose(kd.get_mass_matrix(), D.mass_matrix(s), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(kd.get_bias_forces(), D.inverse_dynamics(s, ds, np.zeros(9)), rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(kd.get_generalized_gravity_forces(), D.inverse_dynamics(s, np.zeros(9), np.zeros(9)),
                               rtol=1e-10, atol=1e-11)
    J = kd.get_frame_jacobian("end_effector_frame")
    assert J.shape == (6, 15)
    l = t["link_names"].index("end_effector_frame")
    np.testing.assert_allclose(J[:, 6:], D.point_jacobian(s, int(t["link_body"][l]), t["link_p"][l]), rtol=1e-10, atol=1e-12)
    H = kd.get_world_transform("end_effector_frame")
    Rw, pw = D.forward_kinematics(s)
    b = int(t["link_body"][l])
    np.testing.assert_allclose(H[:3, 3], pw[b] + Rw[b] @ t["link_p"][l], atol=1e-12)
    np.testing.assert_allclose(H[:3, :3], Rw[b] @ t["link_R"][l], atol=1e-9)
    # centre of mass, momentum (kindyncomputations.py:305-342) and frame bias acceleration (:413-421)
    rc, rv, rm, rg, rJ = D.centroidal(s, ds, t["base_mass"], t["base_mc"])
    np.testing.assert_allclose(kd.get_com_position(), rc, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(kd.get_com_velocity(), rv, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.concatenate(kd.get_momentum()), rm, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(np.concatenate(kd.get_centroidal_momentum()), rg, rtol=1e-10, atol=1e-11)
    Jc = kd.get_com_jacobian()
    assert Jc.shape == (3, 15)
    np.testing.assert_allclose(Jc[:, 6:], rJ, rtol=1e-10, atol=1e-12)
    ref_acc = D.link_motion(s, ds, np.zeros(9), b, t["link_p"][l])
    np.testing.assert_allclose(kd.get_frame_bias_acc("end_effector_frame"), np.r_[ref_acc[2], ref_acc[3]], rtol=1e-9, atol=1e-10)
    # momentum / average-velocity Jacobians (kindyncomputations.py:351-363, 379-427)
    rJ, rl = D.momentum_jacobian(s, t["base_mass"], t["base_mc"], t["base_Io"])
    Jm = kd.get_linear_angular_momentum_jacobian()
    assert Jm.shape == (6, 15)
    np.testing.assert_allclose(Jm[:, 6:], rJ, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(Jm[:3, :3], rl[9] * np.eye(3), atol=1e-12)
    np.testing.assert_allclose(Jm[:, :6], Jm[:, :6].T, atol=1e-12)  # the base block is the locked 6D inertia
    nu = kd.get_model_velocity()
    np.testing.assert_allclose(Jm @ nu, np.concatenate(kd.get_momentum()), rtol=1e-10, atol=1e-11)
    Jg = kd.get_centroidal_total_momentum_jacobian()
    np.testing.assert_allclose(Jg @ nu, rg, rtol=1e-10, atol=1e-11)
    Ja, Jga = kd.get_average_velocity_jacobian(), kd.get_centroidal_average_velocity_jacobian()
    np.testing.assert_allclose(Ja[:, :6], np.eye(6), atol=1e-10)
    np.testing.assert_allclose(Jm[:, :6] @ kd.get_average_velocity(), Jm @ nu, rtol=1e-9, atol=1e-10)
    # centroidal average velocity: linear part = centre-of-mass velocity, and its Jacobian's linear rows = J_com
    np.testing.assert_allclose(kd.get_centroidal_average_velocity()[:3], rv, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(Jga[:3], Jc, rtol=1e-9, atol=1e-11)
    # centre-of-mass bias acceleration against a central difference of J_com(q + h dq) dq
    h = 1e-6
    kd.set_robot_state(s + h * ds, ds); Jp = kd.get_com_jacobian()[:, 6:]
    kd.set_robot_state(s - h * ds, ds); Jn = kd.get_com_jacobian()[:, 6:]
    kd.set_robot_state(s, ds)
    np.testing.assert_allclose(kd.get_com_bias_acc(), (Jp - Jn) @ ds / (2 * h), rtol=1e-5, atol=1e-6)
    # custom joint serialization
    order = list(reversed(kd.joint_serialization()))
    kd2 = KinDynComputations(model_files["panda"], considered_joints=order, world_gravity=np.array([0, 0, -9.806]))
    kd2.set_robot_state(s[::-1], ds[::-1])
    np.testing.assert_allclose(kd2.get_mass_matrix(), D.mass_matrix(s)[::-1, ::-1], rtol=1e-10, atol=1e-12)
    kd.close(); kd2.close()


def test_computed_torque_controller_reference_test(default_world, model_files, oracle):
    """tests/test_scenario/test_custom_controllers.py:20-101 through the scenario API (references inside the joint
    ranges of this Panda model), checked step by step against the oracle and on the reference's own criteria."""
    from scenario import core
    from gym_ignition.context.gazebo import controllers
    gazebo, world = default_world
    assert world.insert_model(model_files["panda"], core.Pose_identity(), "panda")
    panda = world.get_model("panda").to_gazebo()
    assert panda.set_controller_period(gazebo.step_size())
    assert panda.insert_model_plugin(*controllers.ComputedTorqueFixedBase(
        kp=[10.0] * panda.dofs(), ki=[0.0] * panda.dofs(), kd=[3.0] * panda.dofs(), urdf=model_files["panda"],
        joints=panda.joint_names()).args())
    assert all(j.control_mode() == core.JointControlMode_force for j in panda.joints())
    q_ref = [0.0, 0.0, 0.0, -1.5, 0.0, 1.0, 0.0, 0.01, 0.01]
    with pytest.raises(RuntimeError):
        panda.joint_acceleration_targets()
    assert panda.set_joint_position_targets(q_ref)
    assert panda.set_joint_velocity_targets([0.0] * 9)
    assert panda.set_joint_acceleration_targets([0.0] * 9)
    assert panda.joint_acceleration_targets() == pytest.approx([0.0] * 9)
    arm = [j for j in panda.joint_names() if j.startswith("panda_joint")]
    q0 = [q_ref[j] + np.deg2rad(45) * (1 if j != 3 else -0.5) for j in range(7)]
    assert panda.reset_joint_positions(q0, arm) and panda.reset_joint_velocities([0.1] * 7, arm)
    assert gazebo.run(True)
    assert panda.joint_positions(arm) == pytest.approx(q0)

    t, model = oracle.load_urdf(model_files["panda"])
    ref = oracle.Sim(model, 0.001, 1)
    ref.set_controller_period(0.001)
    ref.load_computed_torque([10.0] * 9, [3.0] * 9)
    for j in range(9):
        ref.set_position_target(j, q_ref[j]); ref.set_velocity_target(j, 0.0); ref.set_acceleration_target(j, 0.0)
    for j in range(7):
        ref.reset_position(j, q0[j]); ref.reset_velocity(j, 0.1)
    ref.run(True)
    for k in range(3000):
        assert gazebo.run()
        ref.run(False)
        if k in (0, 10, 200, 1000, 2999):
            got = np.array(panda.joint_positions())
            want = np.array([ref.position(j) for j in range(9)])
            np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-9, err_msg=f"step {k}")
    assert panda.joint_positions() == pytest.approx(panda.joint_position_targets(), abs=np.deg2rad(1))
    assert panda.joint_velocities() == pytest.approx(panda.joint_velocity_targets(), abs=0.05)
    # push on link 4: 100 N for 0.5 s (test_custom_controllers.py:92)
    assert panda.get_link("panda_link4").to_gazebo().apply_world_force([100.0, 0, 0], 0.5)
    l = t["link_names"].index("panda_link4")
    ref.apply_link_wrench(int(t["link_body"][l]), t["link_p"][l], [100.0, 0, 0, 0, 0, 0], 0.5)
    for k in range(4000):
        assert gazebo.run()
        ref.run(False)
        if k in (100, 499, 500, 600):
            np.testing.assert_allclose(panda.joint_positions(), [ref.position(j) for j in range(9)], rtol=1e-6, atol=1e-8)
        if k == 300:
            assert max(abs(a - b) for a, b in zip(panda.joint_positions(), q_ref)) > np.deg2rad(5)
    assert panda.joint_positions() == pytest.approx(panda.joint_position_targets(), abs=np.deg2rad(1))
    assert panda.joint_velocities() == pytest.approx(panda.joint_velocity_targets(), abs=0.05)


def test_link_accelerations_match_oracle(default_world, model_files, oracle):
    """Link world / body accelerations (Link.cpp:206-294) against the oracle, and body = R^T world."""
    from scenario import core
    gazebo, world = default_world
    assert world.insert_model(model_files["panda"])
    panda = world.get_model("panda").to_gazebo()
    assert panda.set_joint_control_mode(core.JointControlMode_force)
    rng = np.random.default_rng(3)
    q0 = list(np.array([0, -0.785, 0, -2.356, 0, 1.571, 0.785, 0.02, 0.02]) + rng.uniform(-0.3, 0.3, 9) * np.r_[np.ones(7), 0, 0])
    assert panda.reset_joint_positions(q0) and panda.reset_joint_velocities(list(rng.uniform(-1, 1, 9)))
    gazebo.run(paused=True)
    for _ in range(5):
        assert panda.set_joint_generalized_force_targets(list(rng.uniform(-3, 3, 9)))
        gazebo.run()
    t, model = oracle.load_urdf(model_files["panda"])
    D = oracle.Dynamics(model)
    q, dq, ddq = np.array(panda.joint_positions()), np.array(panda.joint_velocities()), np.array(panda.joint_accelerations())
    for name in ("panda_link4", "panda_link7", "end_effector_frame", "panda_leftfinger"):
        link = panda.get_link(name)
        l = t["link_names"].index(name)
        ref = D.link_motion(q, dq, ddq, int(t["link_body"][l]), t["link_p"][l])
        np.testing.assert_allclose(link.world_linear_velocity(), ref[0], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(link.world_angular_velocity(), ref[1], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(link.world_linear_acceleration(), ref[2], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(link.world_angular_acceleration(), ref[3], rtol=1e-9, atol=1e-10)
        R = np.array(_quat_R(link.orientation()))
        np.testing.assert_allclose(link.body_linear_acceleration(), R.T @ ref[2], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(link.body_angular_acceleration(), R.T @ ref[3], rtol=1e-8, atol=1e-9)
    assert panda.get_link("panda_link0").world_linear_acceleration() == pytest.approx((0, 0, 0))


def test_joint_friction_setters_and_base_targets(default_world, model_files):
    """Joint.cpp:259-311 (setCoulombFriction / setViscousFriction, only while the model has just been created),
    Model.cpp:1077-1247 (base targets are plain components; getters raise before the first set),
    Link.cpp:529-557 (applyWorldWrenchToCoM)."""
    from scenario import core
    gazebo, world = default_world
    assert world.insert_model(model_files["pendulum"], core.Pose_identity(), "free")
    assert world.insert_model(model_files["pendulum"], core.Pose([2.0, 0, 0], [1.0, 0, 0, 0]), "damped")
    free, damped = world.get_model("free"), world.get_model("damped")
    pivot = damped.get_joint("pivot")
    assert pivot.set_viscous_friction(2.5) and pivot.viscous_friction() == pytest.approx(2.5)
    assert pivot.set_coulomb_friction(0.0) and pivot.coulomb_friction() == 0.0
    for m in (free, damped):
        assert m.reset_joint_positions([1.0]) and m.reset_joint_velocities([0.0])
    # base targets: components that only custom controllers read
    with pytest.raises(RuntimeError):
        free.base_position_target()
    assert free.set_base_position_target([1.0, 2.0, 3.0])
    assert free.base_position_target() == (1.0, 2.0, 3.0) and free.base_orientation_target() == (1.0, 0.0, 0.0, 0.0)
    assert free.set_base_world_velocity_target([0.1, 0, 0], [0, 0, 0.2])
    assert free.base_world_linear_velocity_target() == (0.1, 0.0, 0.0)
    assert free.base_world_angular_velocity_target() == (0.0, 0.0, 0.2)
    with pytest.raises(RuntimeError):
        free.base_world_linear_acceleration_target()
    assert free.set_base_world_linear_acceleration_target([0, 0, 1.0])
    assert free.base_world_linear_acceleration_target() == (0.0, 0.0, 1.0)
    gazebo.run(paused=True)
    for _ in range(300):
        gazebo.run()
    # the damped pendulum lags the undamped one and has lost energy
    assert abs(damped.joint_velocities()[0]) < 0.6 * abs(free.joint_velocities()[0])
    # parameters are frozen once the model has been simulated
    assert not pivot.set_viscous_friction(0.1) and pivot.viscous_friction() == pytest.approx(2.5)
    # a force through the centre of mass equals the same force at the origin plus r x f
    link = free.get_link("pendulum")
    assert link.apply_world_wrench_to_com([0.0, 1.0, 0.0], [0.0, 0.0, 0.0], 0.002)
    assert link.apply_world_wrench_to_co_m([0.0, 1.0, 0.0], [0.0, 0.0, 0.0], 0.0)
    assert gazebo.run()


def test_wrench_through_the_centre_of_mass_equals_wrench_plus_moment_at_the_origin(default_world, model_files):
    """Link.cpp:529-557: applyWorldWrenchToCoM(f, t) = applyWorldWrench(f, t + (W_R_L L_o_com) x f)."""
    from scenario import core
    gazebo, world = default_world
    for k, name in enumerate(("a", "b")):
        assert world.insert_model(model_files["pendulum"], core.Pose([2.0 * k, 0, 0], [1.0, 0, 0, 0]), name)
    a, b = world.get_model("a"), world.get_model("b")
    for m in (a, b):
        assert m.reset_joint_positions([0.7]) and m.reset_joint_velocities([0.0])
    gazebo.run(paused=True)
    la, lb = a.get_link("pendulum"), b.get_link("pendulum")
    f = [0.0, 3.0, 1.0]
    R = np.array(_rot(la.orientation()))
    com = np.array(a._link_com(la._l))
    assert np.linalg.norm(com) > 0.05   # the pendulum's mass sits away from the pivot
    moment = np.cross(R @ com, f)
    assert la.apply_world_wrench_to_com(f, [0.0, 0.0, 0.0], 0.05)
    assert lb.apply_world_wrench(f, list(moment), 0.05)
    for _ in range(10):
        gazebo.run()
    # same joint motion while the link has barely rotated (the moment of b was frozen at the initial orientation)
    assert a.joint_velocities()[0] == pytest.approx(b.joint_velocities()[0], rel=1e-9)
    assert abs(a.joint_velocities()[0]) > 1e-3


def _rot(q):
    w, x, y, z = q
    return [[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]]
