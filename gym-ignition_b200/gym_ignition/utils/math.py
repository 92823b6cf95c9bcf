"""normalize / denormalize front-ends (reference: python/gym_ignition/utils/math.py:10-57)."""
from numbers import Number

import numpy as np
from scenario import gazebo as scenario


def _as_list(v):
    return [v] if isinstance(v, Number) else list(v)


def _unwrap(out):
    return out[0] if len(out) == 1 else np.array(out)


def normalize(input, low, high):
    """Map ``input`` from [low, high] to [-1, 1]; ``None`` bounds leave it untouched."""
    if low is None or high is None:
        return input
    return _unwrap(scenario.normalize(_as_list(input), _as_list(low), _as_list(high)))


def denormalize(input, low, high):
    """Inverse of :py:func:`normalize`."""
    if low is None or high is None:
        return input
    return _unwrap(scenario.denormalize(_as_list(input), _as_list(low), _as_list(high)))
