"""ScenarI/O helpers (reference: python/gym_ignition/utils/scenario.py:13-130)."""
from typing import List, Tuple

import gym.spaces
import gym_ignition_models
import numpy as np
from scenario import gazebo as scenario


def _first_free(base: str, taken) -> str:
    """``base``, ``base1``, ``base2``, ... : the first name that is not in ``taken``."""
    candidate, k = base, 0
    while candidate in taken:
        k += 1
        candidate = f"{base}{k}"
    return candidate


def get_unique_model_name(world, model_name: str) -> str:
    if world.id() == 0:
        raise ValueError("The world is not valid")
    return _first_free(model_name, world.model_names())


def get_unique_world_name(world_name: str) -> str:
    return _first_free(world_name, scenario.ECMSingleton_instance().world_names())


def init_gazebo_sim(step_size: float = 0.001, real_time_factor: float = 1.0, steps_per_run: int = 1,
                    **engine_kwargs) -> Tuple["scenario.GazeboSimulator", "scenario.World"]:
    """Simulator with the default empty world, a ground plane and physics loaded."""
    gazebo = scenario.GazeboSimulator(step_size, real_time_factor, steps_per_run, **engine_kwargs)
    if not gazebo.initialize():
        raise RuntimeError("Failed to initialize Gazebo")
    world = gazebo.get_world()
    if not world.insert_model(gym_ignition_models.get_model_file("ground_plane")):
        raise RuntimeError("Failed to insert the ground plane")
    if not world.set_physics_engine(scenario.PhysicsEngine_dart):
        raise RuntimeError("Failed to insert the physics plugin")
    return gazebo, world


def get_joint_positions_space(model, considered_joints: List[str] = None) -> gym.spaces.Box:
    """Box built from the joint position limits, in the serialisation of ``considered_joints``."""
    names = model.joint_names() if considered_joints is None else considered_joints
    limits = model.joint_limits(names)
    return gym.spaces.Box(low=np.array(limits.min), high=np.array(limits.max))
