"""gym logger bridge (reference: python/gym_ignition/utils/logger.py:39-82)."""
import contextlib

import gym
from gym import logger as _gym_logger
from gym.logger import debug, error, info  # noqa: F401


def warn(msg: str, *args) -> None:
    _gym_logger.warn(msg, *args)


def set_level(level: int) -> None:
    """Set the verbosity of gym and of the ScenarI/O layer together."""
    _gym_logger.set_level(level)
    try:
        from scenario import gazebo as scenario
    except ImportError:
        return
    for threshold, verbosity in ((_gym_logger.DEBUG, scenario.Verbosity_debug), (_gym_logger.INFO, scenario.Verbosity_info),
                                 (_gym_logger.WARN, scenario.Verbosity_warning), (_gym_logger.ERROR, scenario.Verbosity_error)):
        if _gym_logger.MIN_LEVEL <= threshold:
            scenario.set_verbosity(verbosity)
            return
    scenario.set_verbosity(scenario.Verbosity_suppress_all)


@contextlib.contextmanager
def gym_verbosity(level: int):
    previous = gym.logger.MIN_LEVEL
    gym.logger.set_level(level)
    try:
        yield None
    finally:
        gym.logger.set_level(previous)
