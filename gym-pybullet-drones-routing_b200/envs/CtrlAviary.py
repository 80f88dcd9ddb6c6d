"""Batched ``CtrlAviary`` (reference ``envs/CtrlAviary.py``): raw-RPM actions clipped to ``[0, MAX_RPM]``
(``:140``), observation = the 20-float state of every drone (``:117``), dummy reward/flags (``:144-200``)."""
import numpy as np

from ..spaces import Box
from ..utils.enums import DroneModel, Physics
from .BaseAviary import BaseAviary


class CtrlAviary(BaseAviary):
    ENV_KIND = "ctrl"

    def __init__(self, drone_model: DroneModel = DroneModel.CF2X, num_drones: int = 1,
                 neighbourhood_radius: float = np.inf, initial_xyzs=None, initial_rpys=None,
                 physics: Physics = Physics.DYN, pyb_freq: int = 240, ctrl_freq: int = 240, gui=False, record=False,
                 obstacles=False, user_debug_gui=True, output_folder='results', **batch_kwargs):
        super().__init__(drone_model=drone_model, num_drones=num_drones, neighbourhood_radius=neighbourhood_radius,
                         initial_xyzs=initial_xyzs, initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq,
                         ctrl_freq=ctrl_freq, gui=gui, record=record, obstacles=obstacles,
                         user_debug_gui=user_debug_gui, output_folder=output_folder, **batch_kwargs)

    def _actionCode(self):
        return "ctrl_rpm"

    def _actionSpace(self):
        lo = np.zeros((self.NUM_DRONES, 4))
        hi = np.full((self.NUM_DRONES, 4), self.MAX_RPM)
        return Box(low=lo, high=hi, dtype=np.float32)

    def _observationSpace(self):
        inf, pi, mr = np.inf, np.pi, self.MAX_RPM
        lo = np.array([[-inf, -inf, 0., -1., -1., -1., -1., -pi, -pi, -pi, -inf, -inf, -inf, -inf, -inf, -inf, 0., 0., 0., 0.]
                       for _ in range(self.NUM_DRONES)])
        hi = np.array([[inf, inf, inf, 1., 1., 1., 1., pi, pi, pi, inf, inf, inf, inf, inf, inf, mr, mr, mr, mr]
                       for _ in range(self.NUM_DRONES)])
        return Box(low=lo, high=hi, dtype=np.float32)
