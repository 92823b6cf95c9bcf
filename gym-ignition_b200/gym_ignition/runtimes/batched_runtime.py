"""Batched runtime: every env of a B200 simulator advanced by one fused kernel launch per ``step``.

Not in the reference: its runtime steps one world per process (docs/sphinx/info/limitations.rst:19-20). The
relay order of ``GazeboRuntime.step`` / ``reset`` (gazebo_runtime.py:91-140) is preserved per env inside the
kernel: set_action -> run -> observation, reward, done -> (when done) reset_task + paused run.
"""
from typing import Optional

import b2sim
import gym
import numpy as np

from ..base import task as task_mod


class BatchedGazeboRuntime:
    """Vectorised counterpart of GazeboRuntime for tasks that publish a ``batched_spec``.

    ``step(actions)`` takes a CUDA tensor with one action per env and returns ``(obs, reward, done)`` CUDA
    tensors that alias the engine's buffers (zero-copy). Terminated or time-limited envs (``max_episode_steps``,
    the TimeLimit of ``gym.make``) are reset inside the same launch; ``done`` flags them for one step.
    """

    def __init__(self, task_cls: type, num_envs: int, agent_rate: float = 1000.0, physics_rate: float = 1000.0,
                 dtype: str = "float64", device: int = 0, seed: int = 0, env_offset: int = 0,
                 max_episode_steps: int = 5000, **kwargs):
        task = task_cls(agent_rate=agent_rate, **kwargs)
        if not isinstance(task, task_mod.Task):
            raise RuntimeError("The task is not compatible with the runtime")
        env_id = task.batched_spec()
        if env_id is None:
            raise RuntimeError(f"{task_cls.__name__} has no fused kernel (batched_spec() returned None)")
        self.task = task
        self.agent_rate = agent_rate
        self.env = b2sim.BatchedTaskEnv(env_id, num_envs, dtype=dtype, device=device, seed=seed,
                                        env_offset=env_offset, max_episode_steps=max_episode_steps,
                                        physics_rate=physics_rate, agent_rate=agent_rate)
        self.num_envs = num_envs
        # single-env spaces, as the task defines them
        self.action_space, self.observation_space = task.create_spaces()
        task.action_space, task.observation_space = self.action_space, self.observation_space

    def timestamp(self) -> float:
        return self.env.sim.time()

    def reset(self):
        """Fresh episodes for every env; returns the observations, like GazeboRuntime.reset (gazebo_runtime.py:122-140)."""
        return self.env.reset()

    def step(self, actions):
        return self.env.step(actions)

    def step_host(self, actions: np.ndarray, obs: np.ndarray, reward: np.ndarray, done: np.ndarray) -> None:
        self.env.step_host(actions, obs, reward, done)

    def close(self) -> None:
        self.env.close()
