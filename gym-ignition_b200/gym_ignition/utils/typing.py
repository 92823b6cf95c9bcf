"""Type aliases used across the package (reference: python/gym_ignition/utils/typing.py:9-20)."""
from typing import Dict, List, NewType, Tuple, Union

import gym.spaces
import numpy as np

Done = NewType("Done", bool)
Info = NewType("Info", Dict)
Reward = NewType("Reward", float)
Observation = NewType("Observation", np.ndarray)
Action = NewType("Action", Union[np.ndarray, np.number])
SeedList = NewType("SeedList", List[int])
State = NewType("State", Tuple[Observation, Reward, Done, Info])
ActionSpace = NewType("ActionSpace", gym.spaces.Space)
ObservationSpace = NewType("ObservationSpace", gym.spaces.Space)
