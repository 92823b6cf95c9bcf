"""Pins the oracle (and the host-side task mirror) against vectors produced by the REFERENCE's own Python
task classes (tests/golden/make_task_golden.py) and against published known answers."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "task_golden.npz"))
TASK_IDS = [1, 2, 3, 4]


def ulp_diff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 5e-324)


@pytest.mark.parametrize("task", TASK_IDS)
def test_oracle_task_evaluate_matches_reference_tasks(task, oracle):
    states, obs, rew, done = (G[f"task{task}_{k}"] for k in ("states", "obs", "reward", "done"))
    for s, o, r, d in zip(states, obs, rew, done):
        o2, r2, d2 = oracle.task_evaluate(task, s)
        assert d2 == bool(d), f"done differs for state {s}"
        if np.isnan(s).any():
            continue
        # observations: cartpole is a permutation of the state (bit-exact); pendulum goes through libm sin/cos
        assert np.all(ulp_diff(o2, o) <= (1 if task == 1 else 0)), (s, o2, o)
        assert ulp_diff(r2, r) <= 2, (s, r2, r)


@pytest.mark.parametrize("task", TASK_IDS)
def test_oracle_done_is_exact_at_float32_bounds(task):
    """The fixture holds states exactly on, one ulp above and one ulp below the float32-rounded bounds."""
    done = G[f"task{task}_done"]
    assert 0 < done.sum() < len(done)


@pytest.mark.parametrize("task", TASK_IDS)
def test_oracle_action_force_matches_reference_set_action(task, oracle):
    for a, f in zip(G[f"task{task}_actions"], G[f"task{task}_forces"]):
        force, joint = oracle.action_force(task, a)
        assert force == f and joint == 0


@pytest.mark.parametrize("task", TASK_IDS)
def test_oracle_reset_mapping_matches_reference_reset_task(task, oracle):
    for u, st in zip(G[f"task{task}_reset_uniforms"], G[f"task{task}_reset_states"]):
        got = oracle.reset_from_uniforms(task, u)
        if task == 1:
            # q goes through numpy's float32 arctan2 in the reference (a few float32 ulp of error, depending on
            # the numpy build); the oracle and the kernel round a double atan2 to float32 (correctly rounded)
            assert got[1] == st[1] and abs(got[0] - st[0]) <= 4 * np.spacing(np.float32(abs(st[0]))), (u, got, st)
        else:
            assert np.array_equal(got, st), (u, got, st)


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for Philox4x32-10."""
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ---- the host-side mirror of the tasks, driven by the same mock world as the reference was -------------
class _Joint:
    def __init__(self):
        self.q = self.dq = 0.0
        self.last = None
        self.rq = self.rdq = None
        self.mode = None

    def position(self): return self.q
    def velocity(self): return self.dq
    def generalized_force_target(self): return 0.0
    def set_generalized_force_target(self, f): self.last = f; return True
    def set_control_mode(self, m): self.mode = m; return True
    def to_gazebo(self): return self
    def reset(self, q, dq): self.rq, self.rdq = q, dq; return True


class _Model:
    def __init__(self, names): self.j = {n: _Joint() for n in names}
    def get_joint(self, n): return self.j[n]
    def joint_positions(self, names): return [self.j[n].q for n in names]
    def joint_velocities(self, names): return [self.j[n].dq for n in names]
    def to_gazebo(self): return self

    def reset_joint_positions(self, v, names):
        for x, n in zip(v, names): self.j[n].rq = x
        return True

    def reset_joint_velocities(self, v, names):
        for x, n in zip(v, names): self.j[n].rdq = x
        return True


class _World:
    name = "mock"
    def __init__(self, n, m): self.m = {n: m}
    def model_names(self): return list(self.m)
    def get_model(self, n): return self.m[n]


def _mirror_task(task):
    from gym_ignition_environments import tasks
    cls = {1: tasks.pendulum_swingup.PendulumSwingUp, 2: tasks.cartpole_discrete_balancing.CartPoleDiscreteBalancing,
           3: tasks.cartpole_continuous_balancing.CartPoleContinuousBalancing,
           4: tasks.cartpole_continuous_swingup.CartPoleContinuousSwingup}[task]
    t = cls(agent_rate=1000)
    name, joints = ("pendulum", ["pivot"]) if task == 1 else ("cartpole", ["linear", "pivot"])
    model = _Model(joints)
    t.world = _World(name, model)
    t.model_name = name
    t.action_space, t.observation_space = t.create_spaces()
    return t, model


@pytest.mark.parametrize("task", TASK_IDS)
def test_host_task_mirror_matches_reference_tasks(task):
    t, model = _mirror_task(task)
    states, obs, rew, done = (G[f"task{task}_{k}"] for k in ("states", "obs", "reward", "done"))
    for s, o, r, d in zip(states, obs, rew, done):
        if task == 1:
            model.j["pivot"].q, model.j["pivot"].dq = float(s[0]), float(s[1])
        else:
            model.j["linear"].q, model.j["pivot"].q = float(s[0]), float(s[1])
            model.j["linear"].dq, model.j["pivot"].dq = float(s[2]), float(s[3])
        assert np.array_equal(t.get_observation(), o, equal_nan=True)
        assert t.is_done() == bool(d)
        r2 = t.get_reward()
        assert (np.isnan(r2) and np.isnan(r)) or r2 == r
    for a, f in zip(G[f"task{task}_actions"], G[f"task{task}_forces"]):
        t.set_action(int(a) if task == 2 else np.array([a], dtype=np.float32))
        assert model.j["pivot" if task == 1 else "linear"].last == f
    lo_hi = G[f"task{task}_action_low_high"] if task != 1 else np.array([-50.0, 50.0])
    if task != 2:
        assert float(t.action_space.low[0]) == lo_hi[0] and float(t.action_space.high[0]) == lo_hi[1]


# ---- normalisation known answers: tests/test_gym_ignition/test_normalization.py:11-37 -------------------
NORMALIZATION = [
    (1, None, None, 1), (1, 0, None, 1), (1, None, 0, 1),
    (0, -1, 1, 0), (-1, -1, 1, -1), (1, -1, 1, 1),
    ([-1, 0, 1, 2], -2, 2, [-0.5, 0, 0.5, 1]), ([-1, 0, 1, 2], -2., 2., [-0.5, 0, 0.5, 1]),
    ([-1., 0, 1, 2], -2, 2, [-0.5, 0, 0.5, 1]), ([-1., 0, 1, 2], 1, 1, [-1., 0, 1, 2]),
    ([-1, 0, 2.], [-1, -2, 1], [-1, 4, 3], [-1, -0.3333333, 0]),
]


@pytest.mark.parametrize("input,low,high,output", NORMALIZATION)
def test_normalization_known_answers(input, low, high, output):
    from gym_ignition.utils.math import denormalize, normalize
    normalized = normalize(input=input, low=low, high=high)
    assert output == pytest.approx(normalized)
    assert input == pytest.approx(denormalize(input=normalized, low=low, high=high))
