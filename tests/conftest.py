import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__  # noqa: E402

__graft_entry__.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def model_files():
    import gym_ignition_models
    return {n: gym_ignition_models.get_model_file(n) for n in ("pendulum", "cartpole", "panda", "ground_plane")}


@pytest.fixture(scope="session")
def engine_lib():
    """The built CUDA engine; a missing library is a failure, never a skip."""
    import b2sim
    return b2sim._lib.load()
