"""The engine's templated rigid-body code (csrc/b2_rbd.hpp), compiled for the host, against the oracle.

The same templates are instantiated inside the CUDA kernels; running them on the CPU lets the non-GPU suite
check the arithmetic (ABA, RNEA, CRBA, FK, closed-form chain steps) on random states for every model.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HELP = os.path.join(ROOT, "tests", "helpers")


@pytest.fixture(scope="module")
def rbd():
    so = os.path.join(HELP, "librbd_host.so")
    srcs = [os.path.join(HELP, "rbd_host.cpp"), os.path.join(ROOT, "gym-ignition_b200", "csrc", "b2_model.cpp")]
    deps = srcs + [os.path.join(ROOT, "gym-ignition_b200", "csrc", f)
                   for f in ("b2_rbd.hpp", "b2_tree_fast.hpp", "b2_contact.hpp", "b2_model.hpp", "b2_xml.hpp")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so] + srcs)
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.rbd_forward_dynamics.argtypes = [C.c_char_p, dp, dp, C.c_double, dp, dp, dp, dp]
    lib.rbd_forward_dynamics_fast.argtypes = [C.c_char_p, dp, dp, C.c_double, dp, dp, dp, dp]
    lib.rbd_step_fast.argtypes = [C.c_char_p, dp, dp, C.c_double, dp, dp, dp]
    lib.rbd_inverse_dynamics.argtypes = [C.c_char_p, dp, dp, dp, dp, dp, C.c_int, dp]
    lib.rbd_mass_matrix.argtypes = [C.c_char_p, dp, dp, dp, dp]
    lib.rbd_forward_kinematics.argtypes = [C.c_char_p, dp, dp, dp, dp, dp]
    lib.rbd_chain_step.argtypes = [C.c_char_p, dp, dp, C.c_double, dp, dp]
    lib.rbd_centroidal.argtypes = [C.c_char_p] + [dp] * 8
    lib.rbd_momentum.argtypes = [C.c_char_p] + [dp] * 5
    return lib


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


POSES = [((0.0, 0.0, 0.0), (1.0, 0.0, 0.0, 0.0)),
         ((0.3, -0.2, 0.5), (0.9238795325112867, 0.0, 0.3826834323650898, 0.0))]
G = np.array([0.0, 0.0, -9.8])


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "panda"])
@pytest.mark.parametrize("pose", POSES)
def test_dynamics_match_oracle(name, pose, rbd, oracle, model_files):
    xml = open(model_files[name]).read().encode()
    _, model = oracle.load_urdf(model_files[name], base_position=pose[0], base_orientation_wxyz=pose[1])
    D = oracle.Dynamics(model)
    nq = model.nb
    pose7 = np.array(list(pose[0]) + list(pose[1]))
    rng = np.random.default_rng(0)
    for trial in range(10):
        q = np.zeros(16); dq = np.zeros(16); tau = np.zeros(16); ddq_in = np.zeros(16)
        q[:nq] = rng.uniform(-2, 2, nq); dq[:nq] = rng.uniform(-3, 3, nq)
        tau[:nq] = rng.uniform(-10, 10, nq); ddq_in[:nq] = rng.uniform(-5, 5, nq)
        for dt in (0.0, 1e-3):
            out = np.zeros(16)
            assert rbd.rbd_forward_dynamics(xml, dp(pose7), dp(G), dt, dp(q), dp(dq), dp(tau), dp(out)) == nq
            ref = D.forward_dynamics(q[:nq], dq[:nq], tau[:nq], dt)
            np.testing.assert_allclose(out[:nq], ref, rtol=1e-9, atol=1e-9 * (1 + np.abs(ref).max()))
            fast = np.zeros(16)
            assert rbd.rbd_forward_dynamics_fast(xml, dp(pose7), dp(G), dt, dp(q), dp(dq), dp(tau), dp(fast)) == nq
            np.testing.assert_allclose(fast[:nq], ref, rtol=1e-9, atol=1e-9 * (1 + np.abs(ref).max()))
        out = np.zeros(16)
        rbd.rbd_inverse_dynamics(xml, dp(pose7), dp(G), dp(q), dp(dq), dp(ddq_in), 1, dp(out))
        np.testing.assert_allclose(out[:nq], D.inverse_dynamics(q[:nq], dq[:nq], ddq_in[:nq]), rtol=1e-10, atol=1e-10)
        M = np.full(256, np.nan)
        rbd.rbd_mass_matrix(xml, dp(pose7), dp(G), dp(q), dp(M))
        np.testing.assert_allclose(M[:nq * nq].reshape(nq, nq), D.mass_matrix(q[:nq]), rtol=1e-10, atol=1e-12)
        R = np.zeros(16 * 9); p = np.zeros(16 * 3)
        rbd.rbd_forward_kinematics(xml, dp(pose7), dp(G), dp(q), dp(R), dp(p))
        Rr, pr = D.forward_kinematics(q[:nq])
        np.testing.assert_allclose(R[:9 * nq].reshape(nq, 3, 3), Rr, atol=1e-12)
        np.testing.assert_allclose(p[:3 * nq].reshape(nq, 3), pr, atol=1e-12)


@pytest.mark.parametrize("name,kind", [("pendulum", 1), ("cartpole", 2)])
@pytest.mark.parametrize("pose", POSES)
def test_closed_form_chain_step_matches_oracle_step(name, kind, pose, rbd, oracle, model_files):
    """The closed forms the fused kernels evaluate, fitted for (pose, gravity, dt), equal one oracle step
    (DART-style ABA + semi-implicit Euler) to 1e-9 relative."""
    xml = open(model_files[name]).read().encode()
    _, model = oracle.load_urdf(model_files[name], base_position=pose[0], base_orientation_wxyz=pose[1])
    D = oracle.Dynamics(model)
    nq = model.nb
    pose7 = np.array(list(pose[0]) + list(pose[1]))
    rng = np.random.default_rng(1)
    for trial in range(50):
        q = rng.uniform(-3, 3, nq); dq = rng.uniform(-8, 8, nq); tau = rng.uniform(-200, 200, nq)
        if name == "cartpole":
            q[0] = rng.uniform(-2.4, 2.4)  # inside the rail: the closed form carries no joint-limit constraint
        state = np.concatenate([q, dq])
        t = np.zeros(2); t[:nq] = tau
        assert rbd.rbd_chain_step(xml, dp(pose7), dp(G), 1e-3, dp(state), dp(t)) == kind
        q1, dq1, _ = D.step(q, dq, tau, 1e-3)
        np.testing.assert_allclose(state, np.concatenate([q1, dq1]), rtol=1e-9, atol=1e-12)


def test_damped_chain_uses_dart_implicit_damping(rbd, oracle, model_files):
    """Viscous damping enters the joint-space inertia as dt*D (DART's implicit scheme), in both codes."""
    xml = open(model_files["cartpole"]).read().replace('damping="0.0"', 'damping="0.7"')
    t, model = oracle.load_urdf(xml)
    assert t["damping"][0] == 0.7 and t["damping"][1] == 0.7
    D = oracle.Dynamics(model)
    rng = np.random.default_rng(2)
    pose7 = np.array([0, 0, 0, 1.0, 0, 0, 0])
    for trial in range(20):
        q = rng.uniform(-2.4, 2.4, 2); dq = rng.uniform(-8, 8, 2); tau = rng.uniform(-50, 50, 2)
        state = np.concatenate([q, dq])
        assert rbd.rbd_chain_step(xml.encode(), dp(pose7), dp(G), 1e-3, dp(state), dp(tau.copy())) == 2
        q1, dq1, _ = D.step(q, dq, tau, 1e-3)
        np.testing.assert_allclose(state, np.concatenate([q1, dq1]), rtol=1e-9, atol=1e-12)
        # and the implicit term is really there: explicit damping gives a different acceleration
        M = D.mass_matrix(q); h = D.inverse_dynamics(q, dq, np.zeros(2))
        explicit = np.linalg.solve(M, tau - 0.7 * dq - h)
        implicit = np.linalg.solve(M + 1e-3 * 0.7 * np.eye(2), tau - 0.7 * dq - h)
        np.testing.assert_allclose(D.forward_dynamics(q, dq, tau, 1e-3), implicit, rtol=1e-9, atol=1e-10)
        assert np.abs(explicit - implicit).max() > 1e-6


def test_fast_constraint_stage_matches_oracle_step(rbd, oracle, model_files):
    """Joint limits and Coulomb friction through articulated-body impulse responses (the Panda kernel's path)
    against the oracle's dense M^-1 boxed LCP: Panda with fingers on their limits and joint 4 pushed beyond
    its upper limit, and a pendulum with Coulomb friction."""
    pose7 = np.array([0, 0, 0, 1.0, 0, 0, 0])
    xml = open(model_files["panda"]).read().encode()
    _, model = oracle.load_urdf(model_files["panda"])
    D = oracle.Dynamics(model)
    rng = np.random.default_rng(5)
    hit = 0
    for trial in range(30):
        q = np.array([0, -0.785, 0, -2.356, 0, 1.571, 0.785, 0.0, 0.04]) + np.r_[rng.uniform(-0.2, 0.2, 7), 0, 0]
        if trial % 3 == 0:
            q[3] = -0.05          # beyond the upper limit of joint 4 (-0.0698)
        dq = rng.uniform(-1, 1, 9)
        tau = rng.uniform(-5, 5, 9)
        q1, dq1, _ = D.step(q, dq, tau, 1e-3)
        qq, dd, tt = np.zeros(16), np.zeros(16), np.zeros(16)
        qq[:9], dd[:9], tt[:9] = q, dq, tau
        nr = rbd.rbd_step_fast(xml, dp(pose7), dp(G), 1e-3, dp(qq), dp(dd), dp(tt))
        assert nr >= 2
        hit += nr
        np.testing.assert_allclose(dd[:9], dq1, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(qq[:9], q1, rtol=1e-9, atol=1e-12)
    assert hit > 60
    xmlp = open(model_files["pendulum"]).read().replace('friction="0.0"', 'friction="0.05"')
    _, mp = oracle.load_urdf(xmlp)
    Dp = oracle.Dynamics(mp)
    for trial in range(20):
        q, dq, tau = rng.uniform(-3, 3, 1), rng.uniform(-0.5, 0.5, 1), rng.uniform(-0.2, 0.2, 1)
        q1, dq1, _ = Dp.step(q, dq, tau, 1e-3)
        qq, dd, tt = np.zeros(16), np.zeros(16), np.zeros(16)
        qq[0], dd[0], tt[0] = q[0], dq[0], tau[0]
        assert rbd.rbd_step_fast(xmlp.encode(), dp(pose7), dp(G), 1e-3, dp(qq), dp(dd), dp(tt)) == 1
        np.testing.assert_allclose([qq[0], dd[0]], [q1[0], dq1[0]], rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "panda"])
@pytest.mark.parametrize("pose", POSES)
def test_centroidal_quantities_match_oracle(name, pose, rbd, oracle, model_files):
    """Centre of mass, its velocity, momentum about the world origin / the centre of mass and the centre-of-mass
    Jacobian (kindyncomputations.py:305-342): the engine's recursion over body velocities against the oracle's
    Jacobian-based restatement; J_com dq reproduces the centre-of-mass velocity."""
    xml = open(model_files[name]).read().encode()
    t, model = oracle.load_urdf(model_files[name], base_position=pose[0], base_orientation_wxyz=pose[1])
    D = oracle.Dynamics(model)
    nq = model.nb
    pose7 = np.array(list(pose[0]) + list(pose[1]))
    rng = np.random.default_rng(1)
    for trial in range(10):
        q = np.zeros(16); dq = np.zeros(16)
        q[:nq] = rng.uniform(-2, 2, nq); dq[:nq] = rng.uniform(-3, 3, nq)
        com, vel, mom, jac = np.zeros(3), np.zeros(3), np.zeros(12), np.zeros(3 * nq)
        assert rbd.rbd_centroidal(xml, dp(pose7), dp(G), dp(q), dp(dq), dp(com), dp(vel), dp(mom), dp(jac)) == nq
        rc, rv, rm, rg, rJ = D.centroidal(q[:nq], dq[:nq], t["base_mass"], t["base_mc"])
        np.testing.assert_allclose(com, rc, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(vel, rv, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(mom[:6], rm, rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(mom[6:], rg, rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(jac.reshape(3, nq), rJ, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(jac.reshape(3, nq) @ dq[:nq], vel, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "panda"])
@pytest.mark.parametrize("pose", POSES)
def test_momentum_jacobian_matches_oracle(name, pose, rbd, oracle, model_files):
    """Momentum Jacobian and locked inertia (kindyncomputations.py:379-427): the engine's composite-inertia form
    (Ic_j S_j in the base-origin / world-orientation frame) against the oracle's link-by-link restatement through the
    point Jacobians; J dq reproduces the momentum of the centroidal query once moved to the base origin."""
    xml = open(model_files[name]).read().encode()
    t, model = oracle.load_urdf(model_files[name], base_position=pose[0], base_orientation_wxyz=pose[1])
    D = oracle.Dynamics(model)
    nq = model.nb
    pose7 = np.array(list(pose[0]) + list(pose[1]))
    rng = np.random.default_rng(2)
    for trial in range(10):
        q = np.zeros(16); dq = np.zeros(16)
        q[:nq] = rng.uniform(-2, 2, nq); dq[:nq] = rng.uniform(-3, 3, nq)
        jm, locked = np.zeros(6 * nq), np.zeros(10)
        assert rbd.rbd_momentum(xml, dp(pose7), dp(G), dp(q), dp(jm), dp(locked)) == nq
        rJ, rl = D.momentum_jacobian(q[:nq], t["base_mass"], t["base_mc"], t["base_Io"])
        np.testing.assert_allclose(jm.reshape(6, nq), rJ, rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(locked, rl, rtol=1e-11, atol=1e-12)
        _, _, rm, _, _ = D.centroidal(q[:nq], dq[:nq], t["base_mass"], t["base_mc"])
        h = jm.reshape(6, nq) @ dq[:nq]
        np.testing.assert_allclose(h[:3], rm[:3], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(h[3:], rm[3:] - np.cross(np.array(pose[0]), rm[:3]), rtol=1e-10, atol=1e-10)
