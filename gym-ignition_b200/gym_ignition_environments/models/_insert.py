"""Common insertion helper of the model wrappers (reference: python/gym_ignition_environments/models/*.py)."""
from typing import Sequence

import gym_ignition_models
from gym_ignition.utils.scenario import get_unique_model_name
from scenario import core as scenario


def insert_named_model(world, base_name: str, position: Sequence[float], orientation: Sequence[float],
                       model_file: str = None):
    """Insert ``base_name`` (or ``model_file``) under a unique name at the given pose; return the model."""
    name = get_unique_model_name(world, base_name)
    model_file = model_file if model_file is not None else gym_ignition_models.get_model_file(base_name)
    if not world.to_gazebo().insert_model(model_file, scenario.Pose(position, orientation), name):
        raise RuntimeError("Failed to insert model")
    return world.get_model(name)
