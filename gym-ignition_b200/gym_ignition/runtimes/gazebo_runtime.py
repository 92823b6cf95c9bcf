"""Single-env runtime over the ScenarI/O API (reference: python/gym_ignition/runtimes/gazebo_runtime.py).

``step`` relays  set_action -> gazebo.run() -> get_observation -> get_reward -> is_done -> get_info
(gazebo_runtime.py:91-120); ``reset`` relays  reset_task -> paused run -> get_observation (:122-140).
The simulator and the world are created lazily, the world being an empty one with a ground plane and DART
physics (:177-267). Here ``scenario`` is the B200 engine's drop-in module.
"""
import gym_ignition_models
import numpy as np
from scenario import gazebo as scenario

from .. import base, utils
from ..base import runtime
from ..utils import logger


class GazeboRuntime(runtime.Runtime):
    metadata = {"render.modes": ["human"]}

    def __init__(self, task_cls: type, agent_rate: float, physics_rate: float, real_time_factor: float,
                 physics_engine=scenario.PhysicsEngine_dart, world: str = None, **kwargs):
        self._gazebo = None
        self._world = None
        self._physics_rate = physics_rate
        self._real_time_factor = real_time_factor
        self._physics_engine = physics_engine
        self._world_sdf = world
        self._world_name = None

        task = task_cls(agent_rate=agent_rate, **kwargs)
        if not isinstance(task, base.task.Task):
            raise RuntimeError("The task is not compatible with the runtime")
        super().__init__(task=task, agent_rate=agent_rate)

        _ = self.gazebo  # creates the simulator and the world
        self.action_space, self.observation_space = self.task.create_spaces()
        self.task.action_space, self.task.observation_space = self.action_space, self.observation_space
        self.seed()

    # -- Runtime --
    def timestamp(self) -> float:
        return self.world.time()

    # -- gym.Env --
    def step(self, action):
        if not self.action_space.contains(action):
            logger.warn("The action does not belong to the action space")
        self.task.set_action(action)
        assert self.gazebo.run(), "Failed to step gazebo"
        observation = self.task.get_observation()
        assert isinstance(observation, np.ndarray)
        if not self.observation_space.contains(observation):
            logger.warn("The observation does not belong to the observation space")
        reward = self.task.get_reward()
        assert isinstance(reward, float), "Failed to get the reward"
        done = self.task.is_done()
        return observation, reward, done, self.task.get_info()

    def reset(self):
        self.task.reset_task()
        if not self.gazebo.run(paused=True):
            raise RuntimeError("Failed to run Gazebo")
        observation = self.task.get_observation()
        assert isinstance(observation, np.ndarray)
        if not self.observation_space.contains(observation):
            logger.warn("The observation does not belong to the observation space")
        return observation

    def render(self, mode: str = "human", **kwargs) -> None:
        if mode != "human":
            raise ValueError(f"Render mode '{mode}' not supported")
        if not self.gazebo.gui():
            raise RuntimeError("Failed to render the environment")

    def close(self) -> None:
        if not self.gazebo.close():
            raise RuntimeError("Failed to close Gazebo")

    def seed(self, seed: int = None):
        if not self.task.has_world():
            raise RuntimeError("The world has never been created")
        return self.task.seed_task(seed)

    # -- lazily created simulator / world --
    @property
    def gazebo(self) -> "scenario.GazeboSimulator":
        if self._gazebo is not None:
            assert self._gazebo.initialized()
            return self._gazebo
        steps = self._physics_rate / self.agent_rate
        if steps != int(steps):
            logger.warn("Rounding the number of iterations to {} from the nominal {}".format(int(steps), steps))
        self._gazebo = scenario.GazeboSimulator(1.0 / self._physics_rate, self._real_time_factor, int(steps))
        _ = self.world
        assert self._gazebo.initialized()
        return self._gazebo

    @property
    def world(self) -> "scenario.World":
        if self._world is not None:
            assert self.gazebo.initialized()
            return self._world
        if self._gazebo is None:
            raise RuntimeError("Gazebo has not yet been created")
        if self._gazebo.initialized():
            raise RuntimeError("Gazebo was already initialized, cannot insert world")
        if self._world_sdf is None:
            self._world_sdf = ""
            self._world_name = utils.scenario.get_unique_world_name("default")
        else:
            self._world_name = utils.scenario.get_unique_world_name(scenario.get_world_name_from_sdf(self._world_sdf))
        if not self._gazebo.insert_world_from_sdf(self._world_sdf, self._world_name):
            raise RuntimeError("Failed to load SDF world")
        if not self._gazebo.initialize() or not self._gazebo.initialized():
            raise RuntimeError("Failed to initialize Gazebo")
        world = self._gazebo.get_world(self._world_name)
        assert self._world_name in self._gazebo.world_names()
        if self._world_sdf == "":
            if not world.insert_model(gym_ignition_models.get_model_file("ground_plane")):
                raise RuntimeError("Failed to insert the ground plane")
        self.task.world = world
        world.set_physics_engine(engine=self._physics_engine)
        self._world = world
        return self._world
