"""reference: python/gym_ignition_environments/models/panda.py:11-80."""
from typing import List

from gym_ignition.scenario import model_with_file, model_wrapper
from scenario import core as scenario

from ._insert import insert_named_model

#: initial arm configuration (panda.py:42-44)
ARM_Q0 = [0, -0.785, 0, -2.356, 0, 1.571, 0.785]

#: position PID gains at 1 kHz (panda.py:48-58, tests/test_scenario/test_pid_controllers.py:20-30)
PID_GAINS_1000HZ = {
    "panda_joint1": (50, 0, 20),
    "panda_joint2": (10000, 0, 500),
    "panda_joint3": (100, 0, 10),
    "panda_joint4": (1000, 0, 50),
    "panda_joint5": (100, 0, 10),
    "panda_joint6": (100, 0, 10),
    "panda_joint7": (10, 0.5, 0.1),
    "panda_finger_joint1": (100, 0, 50),
    "panda_finger_joint2": (100, 0, 50),
}


class Panda(model_wrapper.ModelWrapper, model_with_file.ModelWithFile):
    def __init__(self, world, position: List[float] = (0.0, 0.0, 0.0), orientation: List[float] = (1.0, 0, 0, 0),
                 model_file: str = None):
        model = insert_named_model(world, "panda", position, orientation, model_file)
        arm = [name for name in model.joint_names() if "panda_joint" in name]
        model.to_gazebo().reset_joint_positions(ARM_Q0, arm)
        if set(model.joint_names()) != set(PID_GAINS_1000HZ):
            raise ValueError("The number of PIDs does not match the number of joints")
        for joint_name, (p, i, d) in PID_GAINS_1000HZ.items():
            if not model.get_joint(joint_name).set_pid(pid=scenario.PID(p, i, d)):
                raise RuntimeError(f"Failed to set the PID of joint '{joint_name}'")
        # the reference sets 1000.0 *seconds* here (panda.py:71); its tests override it with the step size
        assert model.set_controller_period(1000.0)
        super().__init__(model=model)

    @classmethod
    def get_model_file(cls) -> str:
        import gym_ignition_models
        return gym_ignition_models.get_model_file("panda")
