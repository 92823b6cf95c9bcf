"""Minimal stand-in for OpenAI gym (the 0.17-era API gym-ignition was written against, setup.py:46).

Only used when the real ``gym`` package is not installed (it is absent from this image and there is no
network). It implements what the reference's Python layer touches: ``Env``, ``Wrapper``, ``spaces.Box`` /
``Discrete``, ``utils.seeding.np_random``, ``envs.registration.register`` / ``make``, ``wrappers.TimeLimit``
and ``logger``.
"""
from . import logger, spaces, utils, wrappers  # noqa: F401
from .core import Env, Wrapper  # noqa: F401
from . import envs  # noqa: F401
from .envs.registration import make, register, spec  # noqa: F401

__version__ = "0.17.3+b200shim"
