"""Controller plugin contexts (reference: python/gym_ignition/context/gazebo/controllers.py:13-46)."""
from dataclasses import dataclass, field
from typing import Iterable, List, Tuple

from . import plugin

GRAVITY = (0, 0, -9.80665)


@dataclass
class ComputedTorqueFixedBase(plugin.GazeboPlugin):
    """tau = M (ddq_ref - kp (q - q_ref) - kd (dq - dq_ref)) + h, run by the ControllerRunner plugin."""

    urdf: str
    kp: List[float]
    ki: List[float]
    kd: List[float]
    joints: List[str]
    gravity: Tuple[float, float, float] = field(default_factory=lambda: GRAVITY)

    _name: str = field(init=False, repr=False, default="ComputedTorqueFixedBase")
    _plugin_name: str = field(init=False, repr=False, default="ControllerRunner")
    _plugin_class: str = field(init=False, repr=False, default="scenario::plugins::gazebo::ControllerRunner")

    @staticmethod
    def _to_str(values: Iterable) -> str:
        return " ".join(str(v) for v in values)

    def to_xml(self) -> str:
        rows = [("kp", self.kp), ("ki", self.ki), ("kd", self.kd), ("joints", self.joints), ("gravity", self.gravity)]
        body = "".join(f"<{tag}>{self._to_str(v)}</{tag}>" for tag, v in rows)
        return f'<controller name="{self._name}">{body}<urdf>{self.urdf}</urdf></controller>'
