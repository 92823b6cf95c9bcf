"""reference: python/gym_ignition/scenario/model_wrapper.py:9-20."""
import abc

from scenario import core as scenario_core


class ModelWrapper(scenario_core.Model, abc.ABC):
    """A ``scenario.core.Model`` that forwards every attribute it does not define to the wrapped model."""

    def __init__(self, model):
        abc.ABC.__init__(self)
        self.model = model

    def __getattr__(self, name):
        return getattr(self.model, name)
