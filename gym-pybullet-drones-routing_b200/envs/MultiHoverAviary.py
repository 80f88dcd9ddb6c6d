"""Batched ``MultiHoverAviary`` (reference ``envs/MultiHoverAviary.py``): per-drone targets
``INIT_XYZS + [0,0,1/(i+1)]`` (``:71``), summed reward (``:84-88``), terminated (``:101-108``),
truncated (``:121-130``)."""
import numpy as np

from ..utils.enums import ActionType, DroneModel, ObservationType, Physics
from .BaseRLAviary import BaseRLAviary


class MultiHoverAviary(BaseRLAviary):
    ENV_KIND = "multihover"

    def __init__(self, drone_model: DroneModel = DroneModel.CF2X, num_drones: int = 2,
                 neighbourhood_radius: float = np.inf, initial_xyzs=None, initial_rpys=None,
                 physics: Physics = Physics.DYN, pyb_freq: int = 240, ctrl_freq: int = 30, gui=False, record=False,
                 obs: ObservationType = ObservationType.KIN, act: ActionType = ActionType.RPM, **batch_kwargs):
        self.EPISODE_LEN_SEC = 8
        self._num_drones_for_target = num_drones
        super().__init__(drone_model=drone_model, num_drones=num_drones, neighbourhood_radius=neighbourhood_radius,
                         initial_xyzs=initial_xyzs, initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq,
                         ctrl_freq=ctrl_freq, gui=gui, record=record, obs=obs, act=act, **batch_kwargs)

    def _targetPositions(self):
        # MultiHoverAviary.py:71 — with per-env initial poses (E,N,3) every env hovers above ITS OWN start pose
        self.TARGET_POS = self.INIT_XYZS + np.array([[0, 0, 1 / (i + 1)] for i in range(self.NUM_DRONES)])
        return self.TARGET_POS
