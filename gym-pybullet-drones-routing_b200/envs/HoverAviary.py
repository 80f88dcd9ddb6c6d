"""Batched ``HoverAviary`` (reference ``envs/HoverAviary.py``): target ``[0,0,1]``, 8 s episodes (``:51-52``),
reward ``max(0, 2 - ||target - pos||**4)`` (``:77-79``), terminated (``:92-96``), truncated (``:109-117``)."""
import numpy as np

from ..utils.enums import ActionType, DroneModel, ObservationType, Physics
from .BaseRLAviary import BaseRLAviary


class HoverAviary(BaseRLAviary):
    ENV_KIND = "hover"

    def __init__(self, drone_model: DroneModel = DroneModel.CF2X, initial_xyzs=None, initial_rpys=None,
                 physics: Physics = Physics.DYN, pyb_freq: int = 240, ctrl_freq: int = 30, gui=False, record=False,
                 obs: ObservationType = ObservationType.KIN, act: ActionType = ActionType.RPM, **batch_kwargs):
        self.TARGET_POS = np.array([0, 0, 1])
        self.EPISODE_LEN_SEC = 8
        super().__init__(drone_model=drone_model, num_drones=1, initial_xyzs=initial_xyzs, initial_rpys=initial_rpys,
                         physics=physics, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, gui=gui, record=record, obs=obs,
                         act=act, **batch_kwargs)

    def _targetPositions(self):
        return np.asarray(self.TARGET_POS, dtype=np.float64).reshape(1, 3)
