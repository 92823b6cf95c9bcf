"""Property tests of the CPU oracle (hypothesis): identities that hold for ANY state, so the restatement of DART's
step is pinned by more than the handful of hand-picked states of test_oracle_physics.py. SURVEY.md 8(c) oracle plan."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

FLOATS = lambda lo, hi: st.floats(lo, hi, allow_nan=False, allow_infinity=False, width=64)


def vec(n, lo, hi):
    return st.lists(FLOATS(lo, hi), min_size=n, max_size=n).map(np.array)


@pytest.fixture(scope="module")
def panda(oracle, model_files):
    t, model = oracle.load_urdf(model_files["panda"])
    return t, model, oracle.Dynamics(model)


@pytest.fixture(scope="module")
def cartpole(oracle, model_files):
    t, model = oracle.load_urdf(model_files["cartpole"])
    return t, model, oracle.Dynamics(model)


@settings(max_examples=60, deadline=None)
@given(q=vec(9, -2.5, 2.5), dq=vec(9, -3, 3), tau=vec(9, -50, 50))
def test_forward_dynamics_inverts_the_equations_of_motion(panda, q, dq, tau):
    """M(q) ddq + h(q, dq) = tau - D dq with ddq from the articulated-body algorithm (dt = 0: no implicit term)."""
    t, model, D = panda
    damp = np.asarray(t["damping"][:9])
    ddq = D.forward_dynamics(q, dq, tau, 0.0)
    M, h = D.mass_matrix(q), D.inverse_dynamics(q, dq, np.zeros(9))
    scale = max(1.0, np.abs(tau).max(), np.abs(h).max())
    np.testing.assert_allclose(M @ ddq + h, tau - damp * dq, rtol=0, atol=1e-8 * scale)


@settings(max_examples=60, deadline=None)
@given(q=vec(9, -2.5, 2.5), a=vec(9, -1, 1), b=vec(9, -1, 1))
def test_mass_matrix_is_symmetric_positive_definite_and_the_dynamics_are_affine_in_tau(panda, q, a, b):
    t, model, D = panda
    M = D.mass_matrix(q)
    assert np.allclose(M, M.T, atol=1e-12) and np.linalg.eigvalsh(M).min() > 0
    dq = np.zeros(9)
    f0 = D.forward_dynamics(q, dq, np.zeros(9), 0.0)
    fa, fb, fab = (D.forward_dynamics(q, dq, x, 0.0) for x in (a, b, a + b))
    np.testing.assert_allclose((fa - f0) + (fb - f0), fab - f0, rtol=0, atol=1e-7 * max(1.0, np.abs(fab).max()))
    np.testing.assert_allclose(M @ (fa - f0), a, rtol=0, atol=1e-8)      # d(ddq)/d(tau) = M^-1


@settings(max_examples=40, deadline=None)
@given(x=FLOATS(-1, 1), th=FLOATS(-3, 3), dx=FLOATS(-2, 2), dth=FLOATS(-5, 5), f=FLOATS(-200, 200))
def test_implicit_damping_step_solves_its_own_linear_system(cartpole, x, th, dx, dth, f):
    """DART's step: (M + dt D) ddq = tau - D dq - h, then dq += ddq dt, q += dq dt (semi-implicit Euler)."""
    t, model, D = cartpole
    dt = 1e-3
    q, dq, tau = np.array([x, th]), np.array([dx, dth]), np.array([f, 0.0])
    damp = np.asarray(t["damping"][:2])
    q1, dq1, ddq = D.step(q, dq, tau, dt)
    M, h = D.mass_matrix(q), D.inverse_dynamics(q, dq, np.zeros(2))
    np.testing.assert_allclose((M + dt * np.diag(damp)) @ ddq, tau - damp * dq - h, rtol=0,
                               atol=1e-9 * max(1.0, abs(f), np.abs(h).max()))
    np.testing.assert_allclose(dq1, dq + ddq * dt, rtol=0, atol=1e-12 * max(1.0, np.abs(dq1).max()))
    np.testing.assert_allclose(q1, q + dq1 * dt, rtol=0, atol=1e-14)


@settings(max_examples=50, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), env=st.integers(0, 2 ** 40), step=st.integers(0, 2 ** 30),
       task=st.sampled_from([1, 2, 3, 4]))
def test_reset_sampling_is_a_pure_function_inside_the_task_ranges(oracle, seed, env, step, task):
    """Philox-keyed resets (seed, global env index, step): repeatable, independent of batching, inside the ranges of
    the reference tasks (cartpole_*.py reset_task, pendulum_swingup.py:118-127)."""
    a = np.array(oracle.sample_reset(task, seed, env, step))
    b = oracle.sample_reset_batch(task, seed, env, 3, step)[0]
    assert np.array_equal(a, b)
    if task == 1:
        assert -np.pi <= a[0] <= np.pi and abs(a[1]) <= 10.0
    elif task == 4:
        assert abs(a[0]) <= 0.05 and abs(a[2]) <= 0.05 and abs(a[3]) <= 0.05
        assert np.pi - np.deg2rad(60) - 1e-12 <= a[1] <= np.pi + np.deg2rad(60) + 1e-12
    else:
        assert np.abs(a).max() <= 0.05


@settings(max_examples=60, deadline=None)
@given(x=FLOATS(-3, 3), dx=FLOATS(-25, 25), th=FLOATS(-0.5, 0.5), dth=FLOATS(-25, 25))
def test_balancing_done_mask_is_the_float32_box_of_the_reference(oracle, x, dx, th, dth):
    """cartpole_discrete_balancing.py:46-57,111-119: done = observation outside Box(+-[2.4, 20, 12 deg, 1080 deg/s]) whose
    bounds are float32-rounded while the observation stays float64."""
    high = np.array([2.4, 20.0, np.deg2rad(12), np.deg2rad(3 * 360)]).astype(np.float32).astype(np.float64)
    obs, reward, done = oracle.task_evaluate(2, [x, th, dx, dth])
    assert list(obs) == [x, dx, th, dth]
    inside = bool(np.all(np.abs(np.array([x, dx, th, dth])) <= high))
    assert done == (not inside)
    expected = (0.0 if done else 1.0) - 0.1 * abs(x) - 0.1 * abs(dx) - (10.0 if x >= 0.9 * 2.4 else 0.0)
    assert reward == pytest.approx(expected, rel=1e-12, abs=1e-12)
