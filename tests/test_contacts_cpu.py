"""Free bodies with ground / box contacts on the CPU: the oracle's physical known answers
(tests/test_scenario/test_contacts.py:58-236 restated) and the engine's host-compiled templates against it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_rbd_host import dp, rbd  # noqa: F401  (fixture: builds tests/helpers/librbd_host.so)

EDGE, MASS = 0.2, 5.0


def cube_state(z, x=0.0, y=0.0):
    return [x, y, z, 1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0]


def test_cube_drop_contact_force_matches_weight(oracle):
    """5 kg, 0.2 m cube dropped from a 5 cm gap: no contact while falling; after 150 steps it rests on the ground
    with normals (0,0,1) and a total normal force of m g within 0.1 N (test_contacts.py:58-122)."""
    world = oracle.make_world([oracle.make_box_body(MASS, [EDGE] * 3)], [oracle.ground_plane()])
    X = np.array([cube_state(0.15)])
    contacts = oracle.world_step(world, X)
    assert contacts == []
    for _ in range(150):
        contacts = oracle.world_step(world, X)
    assert len(contacts) == 4 and all(c["a"] == 0 and c["b"] == -1 for c in contacts)
    for c in contacts:
        np.testing.assert_allclose(c["n"], [0, 0, 1])
    fz = sum(c["force"][2] for c in contacts)
    assert fz == pytest.approx(MASS * 9.8, abs=0.1)
    assert abs(sum(c["force"][0] for c in contacts)) < 1e-6
    assert X[0, 2] == pytest.approx(EDGE / 2, abs=2e-3) and np.abs(X[0, 7:]).max() < 2e-2


def test_stacked_cubes(oracle):
    """Two cubes stacked (test_contacts.py:125-236): the lower one touches the ground and the upper cube, the
    ground carries both weights, the upper contact carries one."""
    world = oracle.make_world([oracle.make_box_body(MASS, [EDGE] * 3), oracle.make_box_body(MASS, [EDGE] * 3)],
                              [oracle.ground_plane()])
    X = np.array([cube_state(0.15), cube_state(0.4)])
    for _ in range(600):
        contacts = oracle.world_step(world, X)
    pairs = {(c["a"], c["b"]) for c in contacts}
    assert (0, -1) in pairs and ((1, 0) in pairs or (0, 1) in pairs)
    ground = sum(c["force"][2] for c in contacts if c["b"] == -1)
    upper = sum(c["force"][2] for c in contacts if (c["a"], c["b"]) == (1, 0)) - \
        sum(c["force"][2] for c in contacts if (c["a"], c["b"]) == (0, 1))
    assert ground == pytest.approx(2 * MASS * 9.8, abs=1.1)
    assert upper == pytest.approx(MASS * 9.8, abs=1.1)
    assert X[0, 2] == pytest.approx(0.1, abs=3e-3) and X[1, 2] == pytest.approx(0.3, abs=5e-3)


def test_friction_stops_a_sliding_cube(oracle):
    world = oracle.make_world([oracle.make_box_body(MASS, [EDGE] * 3, mu=0.5)], [oracle.ground_plane()])
    X = np.array([cube_state(0.1)])
    X[0, 7] = 1.0                      # 1 m/s along x, Coulomb friction mu = 0.5
    for _ in range(400):
        oracle.world_step(world, X)
    assert abs(X[0, 7]) < 1e-3         # v0 / (mu g) = 0.204 s to stop
    assert X[0, 0] == pytest.approx(1.0 ** 2 / (2 * 0.5 * 9.8), rel=0.08)


def _engine_world_step(rbd, bodies, statics, X, dt=0.001, g=(0, 0, -9.8), iterations=50, erp=0.01, max_erv=1e-3):
    bp = np.array(bodies, float).ravel()
    sp = np.array(statics, float).ravel()
    out = np.zeros(32 * 12)
    gg = np.array(g, float)
    rbd.contact_world_step.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_double,
                                       C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]
    n = rbd.contact_world_step(len(bodies), dp(bp), len(statics), dp(sp), dt, dp(gg), iterations, erp, max_erv, dp(X), dp(out))
    return out[:12 * n].reshape(n, 12)


def test_engine_templates_match_oracle_world_step(rbd, oracle):
    """The engine's templated contact step (csrc/b2_contact.hpp, compiled for the host) against the oracle on a
    tumbling cube hitting the ground and on stacked cubes: states to 1e-9, contact lists identical."""
    I = MASS / 12 * 2 * EDGE ** 2
    body = [MASS, I, 0, 0, 0, I, 0, 0, 0, I, 0, 0, 0, EDGE / 2, EDGE / 2, EDGE / 2, 1.0]
    plane = [3, 0, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1.0]
    world = oracle.make_world([oracle.make_box_body(MASS, [EDGE] * 3), oracle.make_box_body(MASS, [EDGE] * 3)],
                              [oracle.ground_plane()])
    X1 = np.array([cube_state(0.25), cube_state(0.6, 0.02, -0.01)])
    X1[0, 10:] = [2.0, -1.0, 0.5]      # tumbling
    X1[1, 7:10] = [0.1, 0.0, -0.5]
    X2 = X1.copy()
    for step in range(500):
        ref = oracle.world_step(world, X1)
        got = _engine_world_step(rbd, [body, body], [plane], X2)
        assert len(ref) == len(got), step
        np.testing.assert_allclose(X2, X1, rtol=1e-9, atol=1e-10, err_msg=f"step {step}")
        for c, row in zip(ref, got):
            assert (c["a"], c["b"]) == (int(row[0]), int(row[1]))
            np.testing.assert_allclose(row[9:12], c["force"], rtol=1e-6, atol=1e-6)


def test_oracle_external_wrench_on_a_free_body(oracle):
    """b2o_world.ext: a force m g upwards at the root link origin cancels gravity exactly; a torque about a principal axis
    spins the body up at tau / I; a force off the centre of mass adds the moment (o - c) x f."""
    m, edge = 2.0, 0.2
    I = m * edge ** 2 / 6
    world = oracle.make_world([oracle.make_box_body(m, [edge] * 3, inertia=np.eye(3) * I)], [], gravity=(0, 0, -9.8))
    X = np.zeros((1, 13)); X[0, 3] = 1.0; X[0, 2] = 1.0
    world.ext[0][:] = [0, 0, m * 9.8, 0, 0, 0.3]
    for _ in range(100):
        oracle.world_step(world, X)
    assert X[0, 2] == pytest.approx(1.0, abs=1e-12) and abs(X[0, 9]) < 1e-12
    assert X[0, 12] == pytest.approx(0.3 / I * 0.1, rel=1e-12)
    # centre of mass 5 cm along x from the link origin: a force along z at the origin gives a moment -r x f about y
    world = oracle.make_world([oracle.make_box_body(m, [edge] * 3, inertia=np.eye(3) * I, com=(0.05, 0, 0))], [],
                              gravity=(0, 0, 0))
    X = np.zeros((1, 13)); X[0, 3] = 1.0
    world.ext[0][:] = [0, 0, 4.0, 0, 0, 0]
    oracle.world_step(world, X)
    assert X[0, 11] == pytest.approx(0.05 * 4.0 / I * 1e-3, rel=1e-9)      # (o - c) x f = (-0.05, 0, 0) x (0, 0, 4) = (0, 0.2, 0)
