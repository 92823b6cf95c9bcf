"""Coupled worlds on the CPU (BASELINE config 5, examples/panda_pick_and_place.py): the Panda's finger pads against a
cube resting on a table. Checks the oracle's physical behaviour (grasp force, lift) and the engine's host-compiled
coupled step (csrc/b2_contact.hpp coupled_step) against the oracle's independent dense restatement."""
import ctypes as C

import numpy as np
import pytest

from test_rbd_host import dp, rbd  # noqa: F401  (fixture: builds tests/helpers/librbd_host.so)

Q0 = np.array([0, -0.785, 0, -2.356, 0, 1.571, 0.785, 0.04, 0.04])  # models/panda.py:42-44, fingers open
BASE = (0.0, 0.0, 1.0)                                                # examples/panda_pick_and_place.py:217-218 height
CUBE_MASS, EDGE = 0.1, 0.05                                           # "wood cube 5cm"


def pick_scene(oracle, model_files):
    """Panda on a 1 m pedestal, cube between the open finger pads, resting on a table top just below the pads."""
    t, m = oracle.load_urdf(model_files["panda"], base_position=BASE)
    D = oracle.Dynamics(m)
    R, p = D.forward_kinematics(Q0)
    ee = t["link_names"].index("end_effector_frame")
    b = t["link_body"][ee]
    pee = p[b] + R[b] @ t["link_p"][ee]
    zc = round(float(pee[2]), 3)
    cube = oracle.make_box_body(CUBE_MASS, [EDGE] * 3)
    table_centre = [0.307, 0.0, zc - EDGE / 2 - 0.025]
    table = oracle.make_box_static([0.4, 0.4, 0.05], table_centre)
    world = oracle.make_world([cube], [table, oracle.ground_plane()])
    X0 = np.array([[0.307, 0.0, zc, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]])
    return t, m, D, world, oracle.robot_shapes_from_tables(t), X0, table_centre


def test_loaders_agree_on_finger_pads(oracle, model_files):
    """The product loader and the oracle's flattener place the finger pads identically (body, pose, half extents)."""
    import b2sim
    tb = b2sim.ModelInfo.from_file(model_files["panda"]).tables()
    t, _ = oracle.load_urdf(model_files["panda"])
    moving = [k for k in range(tb["nshapes"]) if tb["link_body"][tb["shape_link"][k]] >= 0]
    assert len(moving) == len(t["shapes"]) == 2
    for k, sh in zip(moving, t["shapes"]):
        l = tb["shape_link"][k]
        assert tb["link_body"][l] == sh["body"]
        R = tb["link_R"][l] @ tb["shape_R"][k]
        p = tb["link_R"][l] @ tb["shape_p"][k] + tb["link_p"][l]
        np.testing.assert_allclose(R, sh["R"], atol=1e-12)
        np.testing.assert_allclose(p, sh["p"], atol=1e-12)
        np.testing.assert_allclose(tb["shape_size"][k], sh["size"], atol=1e-12)


def test_oracle_grasp_and_lift(oracle, model_files):
    """ComputedTorqueFixedBase with the gains of examples/panda_pick_and_place.py:34-40: closing the fingers on the
    cube builds up opposite finger contact forces (the example's grasp test reads them, :323-326), and moving the
    arm up lifts the cube with the hand."""
    t, m, D, world, rs, X0, _ = pick_scene(oracle, model_files)
    m.effort[7] = m.effort[8] = 500.0                       # panda_pick_and_place.py:28-32
    sim = oracle.Sim(m, 0.001, 1)
    assert oracle.sim_attach_world(sim, world, rs, X0)
    sim.set_controller_period(0.001)
    for j in range(9):
        sim.reset_position(j, Q0[j])
    sim.run(True)
    sim.load_computed_torque([100.0] * 7 + [10000.0] * 2, [17.5] * 7 + [100.0] * 2)
    for j in range(9):
        sim.set_position_target(j, Q0[j])
        sim.set_velocity_target(j, 0.0)
        sim.set_acceleration_target(j, 0.0)
    for _ in range(200):
        sim.run()
    X = oracle.sim_world_state(sim)
    assert abs(X[0, 2] - X0[0, 2]) < 1e-3 and not any(c["a"] <= -1000 or c["b"] <= -1000 for c in oracle.sim_contacts(sim))
    sim.set_position_target(7, 0.0)                         # close (move_fingers, :138-148)
    sim.set_position_target(8, 0.0)
    for _ in range(400):
        sim.run()
    contacts = oracle.sim_contacts(sim)
    left = sum(-c["force"] if c["a"] == -1000 else c["force"] for c in contacts if -1000 in (c["a"], c["b"]))
    right = sum(-c["force"] if c["a"] == -1001 else c["force"] for c in contacts if -1001 in (c["a"], c["b"]))
    # force ON the cube from each finger: opposite, along the closing direction (world y), tens of newtons
    assert left[1] > 10.0 and right[1] < -10.0
    assert left[1] == pytest.approx(-right[1], rel=0.2)
    assert sim.position(7) == pytest.approx(EDGE / 2, abs=2e-3) and sim.position(8) == pytest.approx(EDGE / 2, abs=2e-3)
    # lift by 6 cm: first-order inverse kinematics of the end-effector height on the arm joints
    q = np.array([sim.position(j) for j in range(9)])
    J = D.point_jacobian(q, 6)[:3, :7]
    dq = np.linalg.pinv(J) @ np.array([0.0, 0.0, 0.06])
    for j in range(7):
        sim.set_position_target(j, q[j] + dq[j])
    for _ in range(1500):
        sim.run()
    X = oracle.sim_world_state(sim)
    assert X[0, 2] - X0[0, 2] > 0.04                        # the cube left the table with the hand
    assert abs(X[0, 1]) < 5e-3                              # and stayed between the pads
    assert not any(c["b"] == -1 for c in oracle.sim_contacts(sim))   # no table contact any more


def _engine_step(rbd, xml, body, statics, robot, q, dq, tau, X):
    rbd.coupled_world_step.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int,
                                       C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_double,
                                       C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double] + [C.POINTER(C.c_double)] * 6
    pose7 = np.array(list(BASE) + [1.0, 0, 0, 0])
    g = np.array([0, 0, -9.8])
    acc, out = np.zeros(16), np.zeros(32 * 12)
    n = rbd.coupled_world_step(xml, dp(pose7), 1, dp(body), 2, dp(statics), 2, dp(robot), 0.001, dp(g), 50, 0.01, 1e-3,
                               dp(q), dp(dq), dp(tau), dp(acc), dp(X), dp(out))
    return n, out[:12 * max(n, 0)].reshape(-1, 12)


def test_engine_coupled_step_matches_oracle(rbd, oracle, model_files):
    """Gravity-compensated arm, 10 N closing force on each finger, then a push on joint 4 that lifts the cube:
    the engine's coupled step (sparse per-body impulses + per-contact joint-space rows) and the oracle's dense
    Jacobian formulation agree to 1e-10 on joint and cube states for 600 steps, with identical contact lists."""
    t, m, D, world, rs, X0, table_centre = pick_scene(oracle, model_files)
    xml = open(model_files["panda"]).read().encode()
    I = CUBE_MASS / 12 * 2 * EDGE ** 2
    body = np.array([CUBE_MASS, I, 0, 0, 0, I, 0, 0, 0, I, 0, 0, 0, EDGE / 2, EDGE / 2, EDGE / 2, 1.0])
    statics = np.array([0, 0.2, 0.2, 0.025, 1, 0, 0, 0, 1, 0, 0, 0, 1] + table_centre + [1.0] +
                       [3, 0, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0], float)
    robot = np.array([7, 0, 0.01, 0.01, 0.015, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0.01, 0.04, 1.0,
                      8, 0, 0.01, 0.01, 0.015, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, -0.01, 0.04, 1.0], float)
    q1, dq1, X1 = Q0.copy(), np.zeros(9), X0.copy()
    q2, dq2, X2 = np.zeros(16), np.zeros(16), X0.copy()
    q2[:9] = Q0
    grasped = False
    for step in range(600):
        tau = D.inverse_dynamics(q1, np.zeros(9), np.zeros(9)) - 5.0 * dq1
        tau[7:] -= 10.0
        if step >= 300:
            tau[3] -= 3.0
        _, ref = oracle.coupled_physics_step(m, world, rs, q1, dq1, tau, X1)
        tau2 = np.zeros(16)
        tau2[:9] = D.inverse_dynamics(q2[:9], np.zeros(9), np.zeros(9)) - 5.0 * dq2[:9]
        tau2[7:9] -= 10.0
        if step >= 300:
            tau2[3] -= 3.0
        n, got = _engine_step(rbd, xml, body, statics, robot, q2, dq2, tau2, X2)
        assert n == len(ref), step
        np.testing.assert_allclose(q2[:9], q1, rtol=0, atol=1e-10, err_msg=f"step {step}")
        np.testing.assert_allclose(dq2[:9], dq1, rtol=0, atol=1e-9, err_msg=f"step {step}")
        np.testing.assert_allclose(X2, X1, rtol=0, atol=1e-9, err_msg=f"step {step}")
        for c, row in zip(ref, got):
            assert (c["a"], c["b"]) == (int(row[0]), int(row[1]))
            np.testing.assert_allclose(row[9:12], c["force"], rtol=1e-6, atol=1e-6)
        grasped = grasped or sum(1 for c in ref if c["a"] <= -1000) == 8
    assert grasped and X1[0, 2] > X0[0, 2] + 5e-4
