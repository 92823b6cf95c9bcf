"""Model files for the B200 engine (stand-in for the pip package ``gym_ignition_models``).

The reference resolves robot descriptions through ``gym_ignition_models.get_model_file``
(/root/reference/python/gym_ignition_environments/models/cartpole.py:47-48,
/root/reference/python/gym_ignition/runtimes/gazebo_runtime.py:249-250). That package is not in the
reference tree, so the model files here are authored for this engine; only the model, link and joint
names are pinned by the reference.
"""
import os
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))


def get_models_path() -> str:
    """Directory that contains one sub-directory per model."""
    return _HERE


def get_model_names() -> List[str]:
    return sorted(d for d in os.listdir(_HERE)
                  if os.path.isdir(os.path.join(_HERE, d)) and not d.startswith("_"))


def get_model_file(robot_name: str) -> str:
    """Absolute path of the URDF (preferred) or SDF file of ``robot_name``."""
    folder = os.path.join(_HERE, robot_name)
    if not os.path.isdir(folder):
        raise FileNotFoundError(f"Model '{robot_name}' not found in {_HERE}")
    for ext in (".urdf", ".sdf"):
        candidate = os.path.join(folder, robot_name + ext)
        if os.path.isfile(candidate):
            return candidate
    raise FileNotFoundError(f"No URDF or SDF file for model '{robot_name}'")


def get_model_string(robot_name: str) -> str:
    with open(get_model_file(robot_name), "r") as f:
        return f.read()


def setup_environment() -> None:
    """The reference uses this to extend IGN_GAZEBO_RESOURCE_PATH; nothing to do here."""
    os.environ.setdefault("B2SIM_MODEL_PATH", _HERE)
