"""reference: python/gym_ignition/runtimes/realtime_runtime.py:18-32 (a NotImplementedError stub there too)."""
from ..base import runtime


class RealTimeRuntime(runtime.Runtime):
    def __init__(self, task_cls: type, robot_cls: type, agent_rate: float, **kwargs):
        raise NotImplementedError

    def timestamp(self) -> float:
        raise NotImplementedError
