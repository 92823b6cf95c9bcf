"""Plugin descriptions turned into ``insert_model_plugin`` arguments
(reference: python/gym_ignition/context/gazebo/plugin.py:17-76)."""
import abc
from dataclasses import dataclass, field
from typing import Tuple

# default namespace of the plugin classes
_SCENARIO_NS = "scenario::plugins::gazebo"


@dataclass
class GazeboPlugin(abc.ABC):
    """A plugin = library name, class name and an XML context; ``args()`` is what
    ``Model.insert_model_plugin(*plugin.args())`` / ``World.insert_world_plugin`` take."""

    _plugin_name: str = field(init=False, repr=False)
    _plugin_class: str = field(init=False, repr=False)

    @abc.abstractmethod
    def to_xml(self) -> str:
        """XML context of the plugin."""

    def args(self) -> Tuple[str, str, str]:
        return str(self._plugin_name), str(self._plugin_class), GazeboPlugin.wrap_in_sdf(self.to_xml())

    @staticmethod
    def wrap_in_sdf(context: str) -> str:
        return f"<sdf version='1.7'>{context}</sdf>"
