#!/usr/bin/env python
"""Generates tests/golden/task_golden.npz by running the REFERENCE's own Python task classes.

Run in the build container only (needs /root/reference); the output is committed and is what the tests on
the GPU box read. The reference task modules
    /root/reference/python/gym_ignition_environments/tasks/{pendulum_swingup,cartpole_*}.py
are imported unmodified (with the reference's gym_ignition.base.task) and driven through a mock ScenarI/O
world that replays given joint states, so the observation / reward / done / set_action / reset_task maths in
the fixture are the reference's, not a restatement. gym itself is not installed in this image; the tasks run
against the gym stand-in of gym-ignition_b200/_shims (0.17 Box.contains / Box.sample semantics).

For resets, the task's RNG is replaced by a stub that returns low + (high - low) * u for a prescribed stream
of uniforms u (numpy's own formula), so that the mapping uniforms -> reset state is captured exactly.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "gym-ignition_b200")
REF = "/root/reference/python"

sys.path[:0] = [REF, PKG, os.path.join(PKG, "_shims"), ROOT]


def load_ref_task(module):
    path = os.path.join(REF, "gym_ignition_environments", "tasks", module + ".py")
    spec = importlib.util.spec_from_file_location("ref_" + module, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class MockJoint:
    def __init__(self):
        self.q = self.dq = 0.0
        self.force_target = 0.0          # what generalized_force_target() returns (0 after a physics step)
        self.last_force_set = None
        self.mode = None
        self.reset_q = self.reset_dq = None

    def position(self): return self.q
    def velocity(self): return self.dq
    def generalized_force_target(self): return self.force_target
    def set_generalized_force_target(self, f): self.last_force_set = f; return True
    def set_control_mode(self, mode): self.mode = mode; return True
    def to_gazebo(self): return self
    def reset(self, q, dq): self.reset_q, self.reset_dq = q, dq; return True


class MockModel:
    def __init__(self, names):
        self.joints = {n: MockJoint() for n in names}

    def get_joint(self, name): return self.joints[name]
    def joint_positions(self, names): return [self.joints[n].q for n in names]
    def joint_velocities(self, names): return [self.joints[n].dq for n in names]
    def to_gazebo(self): return self

    def reset_joint_positions(self, values, names):
        for v, n in zip(values, names): self.joints[n].reset_q = v
        return True

    def reset_joint_velocities(self, values, names):
        for v, n in zip(values, names): self.joints[n].reset_dq = v
        return True


class MockWorld:
    name = "mock"

    def __init__(self, model_name, model):
        self._models = {model_name: model}

    def model_names(self): return list(self._models)
    def get_model(self, name): return self._models[name]


class StreamRNG:
    """uniform(low, high, size) = low + (high - low) * u with u taken from a prescribed stream."""

    def __init__(self, stream):
        self.stream, self.k = list(stream), 0

    def uniform(self, low=0.0, high=1.0, size=None):
        n = int(np.prod(size)) if size is not None else 1
        u = np.array(self.stream[self.k:self.k + n], dtype=np.float64)
        self.k += n
        low, high = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)
        out = low + (high - low) * (u if size is not None else u[0])
        return out

    # Box.sample also calls these for the (empty) unbounded parts of the space
    def normal(self, size=None): return np.zeros(size)
    def exponential(self, size=None): return np.zeros(size)


def make_task(cls, model_name, joint_names):
    task = cls(agent_rate=1000)
    model = MockModel(joint_names)
    task.world = MockWorld(model_name, model)
    task.model_name = model_name
    task.action_space, task.observation_space = task.create_spaces()
    return task, model


def edge_values(threshold):
    t32 = float(np.float32(threshold))
    vals = [t32, np.nextafter(t32, np.inf), np.nextafter(t32, -np.inf), float(threshold), -t32,
            np.nextafter(-t32, -np.inf), np.nextafter(-t32, np.inf), 0.0]
    return [float(v) for v in vals]


def cartpole_states(rng, x_thr, dx_thr, q_thr, dq_thr, n=400):
    st = np.column_stack([rng.uniform(-1.2 * x_thr, 1.2 * x_thr, n), rng.uniform(-1.2 * q_thr, 1.2 * q_thr, n),
                          rng.uniform(-1.2 * dx_thr, 1.2 * dx_thr, n), rng.uniform(-1.2 * dq_thr, 1.2 * dq_thr, n)])
    rows = [st]
    for col, thr in ((0, x_thr), (1, q_thr), (2, dx_thr), (3, dq_thr)):
        for v in edge_values(thr):
            r = np.zeros(4)
            r[col] = v
            rows.append(r[None, :])
    # reward kinks
    for xv in (0.9 * x_thr, np.nextafter(0.9 * x_thr, 0), 0.8 * x_thr, np.nextafter(0.8 * x_thr, 0), x_thr,
               np.nextafter(x_thr, 0)):
        rows.append(np.array([[xv, 0.01, 0.3, -0.2]]))
    rows.append(np.array([[np.nan, 0, 0, 0]]))
    return np.vstack(rows)  # columns: x, q, dx, dq (engine state order)


def run_cartpole(mod_name, cls_name, task_id, out, rng, actions):
    mod = load_ref_task(mod_name)
    cls = getattr(mod, cls_name)
    task, model = make_task(cls, "cartpole", ["linear", "pivot"])
    states = cartpole_states(rng, task._x_threshold, task._dx_threshold, task._q_threshold, task._dq_threshold)
    obs, rew, done = [], [], []
    for x, q, dx, dq in states:
        model.joints["linear"].q, model.joints["linear"].dq = float(x), float(dx)
        model.joints["pivot"].q, model.joints["pivot"].dq = float(q), float(dq)
        obs.append(task.get_observation())
        rew.append(float(task.get_reward()))
        done.append(bool(task.is_done()))
    forces = []
    for a in actions:
        task.set_action(a)
        forces.append(model.joints["linear"].last_force_set)
    # reset mapping
    streams = rng.uniform(0, 1, (64, 4))
    resets = []
    for u in streams:
        task.np_random = StreamRNG(u)
        task.reset_task()
        resets.append([model.joints["linear"].reset_q, model.joints["pivot"].reset_q,
                       model.joints["linear"].reset_dq, model.joints["pivot"].reset_dq])
    assert model.joints["linear"].mode == 2  # JointControlMode_force
    k = f"task{task_id}_"
    out[k + "states"] = states
    out[k + "obs"] = np.array(obs, dtype=np.float64)
    out[k + "reward"] = np.array(rew, dtype=np.float64)
    out[k + "done"] = np.array(done, dtype=np.uint8)
    out[k + "actions"] = np.array([np.asarray(a, dtype=np.float64).ravel()[0] for a in actions])
    out[k + "forces"] = np.array(forces, dtype=np.float64)
    out[k + "reset_uniforms"] = streams
    out[k + "reset_states"] = np.array(resets, dtype=np.float64)
    out[k + "action_low_high"] = np.array([getattr(task.action_space, "low", [0])[0] if hasattr(task.action_space, "low") else 0,
                                           getattr(task.action_space, "high", [1])[0] if hasattr(task.action_space, "high") else 1],
                                          dtype=np.float64)


def run_pendulum(out, rng):
    mod = load_ref_task("pendulum_swingup")
    task, model = make_task(mod.PendulumSwingUp, "pendulum", ["pivot"])
    n = 400
    states = np.column_stack([rng.uniform(-7, 7, n), rng.uniform(-12, 12, n)])
    extra = [[0.3, v] for v in edge_values(10.0)] + [[np.pi, 0.0], [-np.pi / 2, 9.99], [np.nan, 0.0], [0.0, np.nan]]
    states = np.vstack([states, np.array(extra)])
    obs, rew, done = [], [], []
    pivot = model.joints["pivot"]
    for q, dq in states:
        pivot.q, pivot.dq = float(q), float(dq)
        obs.append(task.get_observation())
        rew.append(float(task.get_reward()))
        done.append(bool(task.is_done()))
    actions = [np.array([v], dtype=np.float32) for v in (-50, -3.25, 0, 12.125, 50)]
    forces = []
    for a in actions:
        task.set_action(a)
        forces.append(pivot.last_force_set)
    streams = rng.uniform(0, 1, (64, 3))
    resets = []
    for u in streams:
        task.observation_space.np_random = StreamRNG(u)
        task.reset_task()
        resets.append([pivot.reset_q, pivot.reset_dq])
    out["task1_states"] = states
    out["task1_obs"] = np.array(obs, dtype=np.float64)
    out["task1_reward"] = np.array(rew, dtype=np.float64)
    out["task1_done"] = np.array(done, dtype=np.uint8)
    out["task1_actions"] = np.array([float(a[0]) for a in actions])
    out["task1_forces"] = np.array(forces, dtype=np.float64)
    out["task1_reset_uniforms"] = streams
    out["task1_reset_states"] = np.array(resets, dtype=np.float64)


def main():
    rng = np.random.default_rng(20201018)
    out = {}
    run_pendulum(out, rng)
    cont = lambda vals: [np.array([v], dtype=np.float32) for v in vals]
    run_cartpole("cartpole_discrete_balancing", "CartPoleDiscreteBalancing", 2, out, rng, [0, 1, np.int64(1), 0])
    run_cartpole("cartpole_continuous_balancing", "CartPoleContinuousBalancing", 3, out, rng, cont((-50, -1.5, 0, 33.3, 50)))
    run_cartpole("cartpole_continuous_swingup", "CartPoleContinuousSwingup", 4, out, rng, cont((-200, -7.7, 0, 123.456, 200)))
    path = os.path.join(HERE, "task_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
