"""Pins the oracle's physics and ScenarI/O bookkeeping against the physical known answers the reference's
own tests use and against closed-form mechanics. Parity vs DART trajectories is unpinned
(no golden trajectories exist in the reference); these are the independent checks that remain."""
import ctypes as C

import numpy as np
import pytest

MODE_IDLE, MODE_FORCE, MODE_VELOCITY, MODE_FOLLOWER, MODE_POSITION = 1, 2, 3, 4, 5
DBL_MAX = np.finfo(np.float64).max


def test_pendulum_matches_closed_form_equation(oracle, model_files):
    """tests/.python/test_pendulum_wrt_ground_truth.py:19-67,116-180: m=1, L=0.5, r=0.01,
    I = m (4 L^2 + 3 r^2) / 12, theta'' = (m g L/2 sin(theta) + tau) / I, semi-implicit Euler @ 4 kHz,
    simulation vs equation within 3 degrees."""
    g = 9.8182
    _, model = oracle.load_urdf(model_files["pendulum"], gravity=(0, 0, -g))
    D = oracle.Dynamics(model)
    m, L, r = 1.0, 0.5, 0.01
    I = m * (4 * L ** 2 + 3 * r ** 2) / 12
    dt = 1.0 / 4000
    rng = np.random.default_rng(0)
    q, dq = np.array([np.deg2rad(10.0)]), np.array([0.0])
    th, dth = q[0], 0.0
    for k in range(4000):
        tau = rng.uniform(-0.5, 0.5)
        q, dq, _ = D.step(q, dq, [tau], dt)
        ddth = (m * g * L / 2 * np.sin(th) + tau) / I
        dth += ddth * dt
        th += dth * dt
    assert abs(q[0] - th) < 1e-9 * max(1.0, abs(th))      # same integrator, same equation: far inside 3 degrees
    assert abs(np.rad2deg(q[0] - th)) < 3.0


def test_cartpole_matches_textbook_equations(oracle, model_files):
    """Lagrangian cart-pole with the URDF's parameters (pole COM at l = 0.5, inertia about COM)."""
    t, model = oracle.load_urdf(model_files["cartpole"])
    D = oracle.Dynamics(model)
    mc, mp, l, Ic, g = 1.0, 0.1, 0.5, 0.008349, 9.8
    rng = np.random.default_rng(1)
    for _ in range(20):
        x, th, dx, dth = rng.uniform(-1, 1), rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(-4, 4)
        f, tau = rng.uniform(-50, 50), rng.uniform(-1, 1)
        # pole rotates about +y from the upright: COM at (l sin th, l cos th)
        M = np.array([[mc + mp, mp * l * np.cos(th)], [mp * l * np.cos(th), Ic + mp * l * l]])
        h = np.array([-mp * l * np.sin(th) * dth ** 2, -mp * g * l * np.sin(th)])
        ref = np.linalg.solve(M, np.array([f, tau]) - h)
        got = D.forward_dynamics([x, th], [dx, dth], [f, tau], 0.0)
        np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "panda"])
def test_aba_equals_crba_rnea(name, oracle, model_files):
    """M(q) ddq + h(q, dq) = tau - D dq for the articulated-body solution (dt = 0: no implicit term)."""
    t, model = oracle.load_urdf(model_files[name])
    D = oracle.Dynamics(model)
    nq = model.nb
    damp = np.asarray(t["damping"][:nq])
    rng = np.random.default_rng(2)
    for _ in range(10):
        q, dq, tau = rng.uniform(-2, 2, nq), rng.uniform(-2, 2, nq), rng.uniform(-10, 10, nq)
        ddq = D.forward_dynamics(q, dq, tau, 0.0)
        M, h = D.mass_matrix(q), D.inverse_dynamics(q, dq, np.zeros(nq))
        np.testing.assert_allclose(M @ ddq + h, tau - damp * dq, rtol=1e-9, atol=1e-9)
        assert np.allclose(M, M.T, atol=1e-12) and np.all(np.linalg.eigvalsh(M) > 0)


def test_energy_is_conserved_by_the_undamped_cartpole(oracle, model_files):
    _, model = oracle.load_urdf(model_files["cartpole"])
    D = oracle.Dynamics(model)
    q, dq = np.array([0.0, 2.5]), np.array([0.3, -0.5])
    e0 = D.energy(q, dq)
    drift = 0.0
    for _ in range(3000):
        q, dq, _ = D.step(q, dq, [0.0, 0.0], 1e-3)
        drift = max(drift, abs(D.energy(q, dq) - e0))
    assert drift < 2e-3 * max(1.0, abs(e0))  # symplectic Euler: bounded O(dt) oscillation, no secular growth


def test_jacobian_is_the_derivative_of_forward_kinematics(oracle, model_files):
    t, model = oracle.load_urdf(model_files["panda"])
    D = oracle.Dynamics(model)
    rng = np.random.default_rng(3)
    q = rng.uniform(-1, 1, 9)
    body, pt = 6, np.array([0.01, -0.02, 0.1])
    J = D.point_jacobian(q, body, pt)
    eps = 1e-6
    for j in range(9):
        dqv = np.zeros(9); dqv[j] = eps
        Rp, pp = D.forward_kinematics(q + dqv)
        Rm, pm = D.forward_kinematics(q - dqv)
        lin = ((pp[body] + Rp[body] @ pt) - (pm[body] + Rm[body] @ pt)) / (2 * eps)
        np.testing.assert_allclose(J[:3, j], lin, atol=1e-8)


def test_pid_known_answer(oracle):
    """ignition::math::PID::Update hand computation: p=2, i=0.5, d=0.1, limits +-3, dt=0.01."""
    from oracle.oracle import lib
    buf = (C.c_double * 13)()
    lib().b2o_pid_init(buf, 2.0, 0.5, 0.1, 1.0, -1.0, 3.0, -3.0, 0.25)
    cmd1 = lib().b2o_pid_update(buf, 0.4, 0.01)
    # iErr = 0.5*0.01*0.4 = 0.002 ; dErr = (0.4-0)/0.01 = 40 ; cmd = 0.25 - 0.8 - 0.002 - 4.0 -> clamp -3
    assert cmd1 == -3.0
    cmd2 = lib().b2o_pid_update(buf, 0.4, 0.01)
    # iErr = 0.004 ; dErr = 0 ; cmd = 0.25 - 0.8 - 0.004
    assert cmd2 == pytest.approx(0.25 - 0.8 - 0.004, abs=1e-15)
    assert lib().b2o_pid_update(buf, float("nan"), 0.01) == 0.0
    assert lib().b2o_pid_update(buf, 0.4, 0.0) == 0.0


def make_sim(oracle, model_files, name, dt=0.001, steps=1, **kw):
    _, model = oracle.load_urdf(model_files[name], **kw)
    return oracle.Sim(model, dt, steps), model


def test_time_bookkeeping(oracle, model_files):
    """tests/test_scenario/test_world.py:149-219: time = n dt exactly, paused runs do not advance it."""
    for dt in (0.001, 1e-9, 0.25):
        sim, _ = make_sim(oracle, model_files, "pendulum", dt=dt)
        assert sim.time() == 0.0
        sim.run(True)
        assert sim.time() == 0.0
        sim.run(False)
        assert sim.time() == pytest.approx(dt, rel=1e-12)
        sim.run(False)
        assert sim.time() == pytest.approx(2 * dt, rel=1e-12)
    with pytest.raises(ValueError):
        oracle.Sim(make_sim(oracle, model_files, "pendulum")[1], 0.0, 1)
    with pytest.raises(ValueError):
        oracle.Sim(make_sim(oracle, model_files, "pendulum")[1], 0.001, 0)


def test_resets_are_deferred_to_the_next_run(oracle, model_files):
    """tests/test_scenario/test_model.py:69-110."""
    sim, _ = make_sim(oracle, model_files, "cartpole")
    sim.reset_position(0, 0.3); sim.reset_velocity(1, -1.5)
    assert sim.position(0) == 0.0 and sim.velocity(1) == 0.0
    sim.run(True)
    assert sim.position(0) == 0.3 and sim.velocity(1) == -1.5 and sim.time() == 0.0


def test_force_command_is_one_shot(oracle, model_files):
    """tests/.python/test_joint_force.py:9-81 + Physics.cpp:2250-2254: consumed by ONE physics iteration."""
    sim1, _ = make_sim(oracle, model_files, "cartpole", steps=2)
    sim2, _ = make_sim(oracle, model_files, "cartpole", steps=1)
    for s in (sim1, sim2):
        s.set_control_mode(0, MODE_FORCE)
        assert s.force_target(0) == 0.0
        assert s.set_force_target(0, 10.0)
    sim1.run(False)               # 2 iterations: force in the first only
    sim2.run(False); sim2.run(False)
    assert sim1.position(0) == sim2.position(0) and sim1.velocity(0) == sim2.velocity(0)
    assert sim1.force_target(0) == 0.0
    assert 0.009 < sim1.velocity(0) < 0.011   # one iteration of 10 N on ~1 kg, then coasting


def test_control_mode_semantics(oracle, model_files):
    """Joint.cpp:369-460: switching deletes targets, Position seeds the target with the current position,
    reading an unset target raises; PositionInterpolated is rejected."""
    sim, _ = make_sim(oracle, model_files, "pendulum")
    sim.reset_position(0, 0.7); sim.run(True)
    with pytest.raises(RuntimeError):
        sim.position_target(0)
    assert not sim.set_force_target(0, 1.0)           # Idle does not accept force targets
    assert sim.set_control_mode(0, MODE_POSITION)
    assert sim.position_target(0) == 0.7
    with pytest.raises(RuntimeError):
        sim.force_target(0)
    assert not sim.set_velocity_target(0, 1.0)
    assert sim.set_control_mode(0, MODE_FORCE)
    with pytest.raises(RuntimeError):
        sim.position_target(0)
    assert not sim.set_control_mode(0, 6)


def test_pid_rate_gate(oracle, model_files):
    """JointController.cpp:128-169: the first iteration always computes; afterwards only when the elapsed
    time reaches the controller period; in between the last command is re-applied."""
    sim, _ = make_sim(oracle, model_files, "pendulum")
    sim.set_controller_period(0.003)
    sim.set_pid(0, 10.0, 0.0, 0.0, DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX, 0.0)
    sim.set_control_mode(0, MODE_POSITION)
    sim.set_position_target(0, 0.5)
    from oracle.oracle import lib
    cmds = []
    for k in range(7):
        sim.run(False)
        # the PID command that was applied on this step is pid.cmd; recover it from the error history
        cmds.append(None)
    # replay by hand: compute at steps 1, 4, 7 (elapsed 3 dt), hold in between
    ref, _ = make_sim(oracle, model_files, "pendulum")
    ref.set_control_mode(0, MODE_FORCE)
    cmd = 0.0
    for k in range(1, 8):
        if k in (1, 4, 7):
            cmd = -10.0 * (ref.position(0) - 0.5)
        ref.set_force_target(0, cmd)
        ref.run(False)
    assert sim.position(0) == pytest.approx(ref.position(0), rel=1e-13, abs=1e-16)
    assert sim.velocity(0) == pytest.approx(ref.velocity(0), rel=1e-13, abs=1e-16)


def test_panda_position_pid_holds_and_tracks(oracle, model_files):
    """tests/test_scenario/test_pid_controllers.py:33-115 restated on the oracle: Panda inserted at q = 0 with
    joint1 / joint6 at mid-range, gains of :20-30, controller period = dt; holds within 1 degree for 1000
    steps, then tracks 0.33 Hz sines of 0.9 * range / 2 on joint1 and joint6 within 3 degrees."""
    from gym_ignition_environments.models.panda import PID_GAINS_1000HZ
    sim, model = make_sim(oracle, model_files, "panda")
    t, _ = oracle.load_urdf(model_files["panda"])
    names = t["joint_names"]
    lo, hi = np.asarray(t["lower"]), np.asarray(t["upper"])
    j1, j6 = names.index("panda_joint1"), names.index("panda_joint6")
    rng1, rng6 = abs(hi[j1] - lo[j1]), abs(hi[j6] - lo[j6])
    sim.reset_position(j1, lo[j1] + rng1 / 2)
    sim.reset_position(j6, lo[j6] + rng6 / 2)
    sim.run(True)
    sim.set_controller_period(0.001)
    for j, n in enumerate(names):
        p, i, d = PID_GAINS_1000HZ[n]
        sim.set_pid(j, p, i, d, DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX, 0.0)
        assert sim.set_control_mode(j, MODE_POSITION)
    targets = [sim.position_target(j) for j in range(9)]
    assert targets == [sim.position(j) for j in range(9)]
    for _ in range(1000):
        sim.run(False)
    assert max(abs(sim.position(j) - targets[j]) for j in range(9)) < np.deg2rad(1.0)
    q01, q06 = sim.position(j1), sim.position(j6)
    for k in range(5000):
        s = np.sin(2 * np.pi * 0.33 * k * 0.001)
        r1, r6 = q01 + 0.9 * rng1 / 2 * s, q06 + 0.9 * rng6 / 2 * s
        sim.set_position_target(j1, r1)
        sim.set_position_target(j6, r6)
        sim.run(False)
        assert abs(sim.position(j1) - r1) < np.deg2rad(3.0), k
        assert abs(sim.position(j6) - r6) < np.deg2rad(3.0), k


PENDULUM_FRICTION = None


def test_velocity_follower_and_friction_settling(oracle, model_files):
    """tests/test_scenario/test_velocity_direct.py:19-82: VelocityFollowerDart reaches the target velocity in
    ONE step; with Coulomb 0.01 + viscous 0.2 a pendulum released at 90 degrees settles at the bottom."""
    xml = open(model_files["pendulum"]).read().replace('damping="0.0" friction="0.0"', 'damping="0.2" friction="0.01"')
    t, model = oracle.load_urdf(xml)
    assert t["damping"][0] == 0.2 and t["friction"][0] == 0.01
    sim = oracle.Sim(model, 0.001, 1)
    sim.reset_position(0, np.deg2rad(90)); sim.run(True)
    for _ in range(5000):
        sim.run(False)
    # q = 0 is upright in this model: the stable equilibrium is at +-180 degrees
    assert abs(abs(np.rad2deg(sim.position(0))) - 180.0) < 0.3 or abs(np.rad2deg(sim.position(0)) % 360 - 180) < 0.3
    sim2 = oracle.Sim(model, 0.001, 1)
    assert sim2.set_control_mode(0, MODE_FOLLOWER)
    sim2.set_velocity_target(0, np.pi)
    sim2.run(False)
    assert sim2.velocity(0) == pytest.approx(np.pi, rel=1e-6)
    sim2.set_velocity_target(0, -np.pi)
    sim2.run(False)
    assert sim2.velocity(0) == pytest.approx(-np.pi, rel=1e-6)


def test_joint_limits_stop_the_cart(oracle, model_files):
    """ign-physics enforces SDF position limits: the cart pushed against the end of the rail stops there."""
    sim, _ = make_sim(oracle, model_files, "cartpole")
    sim.set_control_mode(0, MODE_FORCE)
    for _ in range(3000):
        sim.set_force_target(0, 30.0)
        sim.run(False)
    assert 2.6 <= sim.position(0) < 2.65
    assert abs(sim.velocity(0)) < 1e-9


def test_computed_torque_controller_reaches_references_and_recovers(oracle, model_files):
    """tests/test_scenario/test_custom_controllers.py:20-101 restated on the oracle: ComputedTorqueFixedBase
    (kp 10, kd 3, period = dt) brings the Panda from 45 degrees / 0.1 rad/s to the references within 1 degree and
    0.05 rad/s in 3000 steps, and again after a 100 N, 0.5 s push on panda_link4. The references are placed inside
    the joint ranges (the reference test uses q = 0, which this model's joint 4 limit does not allow)."""
    t, model = make_sim(oracle, model_files, "panda")[1], None
    t, model = oracle.load_urdf(model_files["panda"])
    sim = oracle.Sim(model, 0.001, 1)
    assert sim.set_controller_period(0.001)
    assert sim.load_computed_torque([10.0] * 9, [3.0] * 9)
    q_ref = np.array([0.0, 0.0, 0.0, -1.5, 0.0, 1.0, 0.0, 0.01, 0.01])
    sim.run(False)                              # references not set yet: "the controller is not stepping"
    assert sim.velocity(1) != 0.0               # the arm just falls under gravity
    for j in range(9):
        assert sim.set_position_target(j, q_ref[j]) and sim.set_velocity_target(j, 0.0)
        assert sim.set_acceleration_target(j, 0.0)
    for j in range(7):
        sim.reset_position(j, q_ref[j] + np.deg2rad(45) * (1 if j != 3 else -0.5))
        sim.reset_velocity(j, 0.1)
    sim.run(True)
    for _ in range(3000):
        sim.run(False)
    q = np.array([sim.position(j) for j in range(9)]); dq = np.array([sim.velocity(j) for j in range(9)])
    assert np.abs(q - q_ref).max() < np.deg2rad(1) and np.abs(dq).max() < 0.05
    l = t["link_names"].index("panda_link4")
    assert sim.apply_link_wrench(int(t["link_body"][l]), t["link_p"][l], [100.0, 0, 0, 0, 0, 0], 0.5)
    for _ in range(300):
        sim.run(False)
    pushed = np.abs(np.array([sim.position(j) for j in range(9)]) - q_ref).max()
    assert pushed > np.deg2rad(5)                # the push really displaced the arm
    for _ in range(3700):
        sim.run(False)
    q = np.array([sim.position(j) for j in range(9)]); dq = np.array([sim.velocity(j) for j in range(9)])
    assert np.abs(q - q_ref).max() < np.deg2rad(1) and np.abs(dq).max() < 0.05


def test_link_wrench_duration_counts_iterations(oracle, model_files):
    """helpers.h:300-345: a wrench of duration D applied at t0 acts on every iteration whose post-step time has
    not yet reached t0 + D, i.e. ceil(D / dt) iterations, at least one."""
    for duration, expected in ((0.0, 1), (0.0035, 4), (0.002, 2)):
        t, model = oracle.load_urdf(model_files["pendulum"], gravity=(0, 0, 0))
        sim = oracle.Sim(model, 0.001, 1)
        sim.set_control_mode(0, MODE_FORCE)
        l = t["link_names"].index("pendulum")
        sim.apply_link_wrench(int(t["link_body"][l]), t["link_p"][l], [0, 0, 0, 1.0, 0, 0], duration)
        speeds = []
        for _ in range(8):
            sim.run(False)
            speeds.append(sim.velocity(0))
        changes = np.count_nonzero(np.abs(np.diff([0.0] + speeds)) > 1e-12)
        assert changes == expected, (duration, speeds)


def test_link_acceleration_is_the_derivative_of_link_velocity(oracle, model_files):
    """tests/test_scenario/test_link_velocities.py:86-318: link linear / angular velocities and accelerations are
    consistent with finite differences (the reference checks abs 1e-2 / 5e-3 and 0.5 / 0.2 at 10 kHz)."""
    t, model = oracle.load_urdf(model_files["panda"])
    D = oracle.Dynamics(model)
    rng = np.random.default_rng(7)
    l = t["link_names"].index("panda_link7")
    body, pt = int(t["link_body"][l]), t["link_p"][l]
    q, dq = rng.uniform(-1, 1, 9) + np.array([0, -0.785, 0, -2.356, 0, 1.571, 0.785, 0.02, 0.02]), rng.uniform(-1, 1, 9)
    dt = 1e-4
    prev = None
    for step in range(50):
        tau = rng.uniform(-5, 5, 9)
        q, dq, ddq = D.step(q, dq, tau, dt)
        Rw, pw = D.forward_kinematics(q)
        now = D.link_motion(q, dq, ddq, body, pt)
        pos = pw[body] + Rw[body] @ pt
        if prev is not None:
            np.testing.assert_allclose((pos - prev[0]) / dt, now[0], atol=1e-2)        # velocity vs d(position)/dt
            np.testing.assert_allclose((now[0] - prev[1][0]) / dt, now[2], atol=0.5)   # linear acceleration
            np.testing.assert_allclose((now[1] - prev[1][1]) / dt, now[3], atol=0.2)   # angular acceleration
        prev = (pos, now)
