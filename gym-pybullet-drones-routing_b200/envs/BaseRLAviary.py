"""Batched ``BaseRLAviary`` (reference ``envs/BaseRLAviary.py``): action ring of ``ctrl_freq//2`` entries
(``:66-67``), action spaces (``:132-156``), ``_preprocessAction`` for RPM/PID/VEL/ONE_D_RPM/ONE_D_PID
(``:160-239``) and the KIN observation ``[pos rpy vel ang_v | ring]`` (``:307-319``) — all inside ``gpd_step``."""
import numpy as np

from ..spaces import Box
from ..utils.enums import ActionType, DroneModel, ObservationType, Physics
from .BaseAviary import BaseAviary


class BaseRLAviary(BaseAviary):
    def __init__(self, drone_model: DroneModel = DroneModel.CF2X, num_drones: int = 1,
                 neighbourhood_radius: float = np.inf, initial_xyzs=None, initial_rpys=None,
                 physics: Physics = Physics.DYN, pyb_freq: int = 240, ctrl_freq: int = 240, gui=False, record=False,
                 obs: ObservationType = ObservationType.KIN, act: ActionType = ActionType.RPM, **batch_kwargs):
        self.ACTION_BUFFER_SIZE = int(ctrl_freq // 2)
        if obs != ObservationType.KIN:
            raise NotImplementedError("ObservationType.RGB needs the PyBullet renderer: out of scope")
        self.OBS_TYPE = obs
        self.ACT_TYPE = act
        if act in [ActionType.PID, ActionType.VEL, ActionType.ONE_D_PID] and drone_model not in [DroneModel.CF2X, DroneModel.CF2P]:
            raise ValueError("[ERROR] in BaseRLAviary.__init()__, no controller is available for the specified drone_model")
        super().__init__(drone_model=drone_model, num_drones=num_drones, neighbourhood_radius=neighbourhood_radius,
                         initial_xyzs=initial_xyzs, initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq,
                         ctrl_freq=ctrl_freq, gui=gui, record=record, obstacles=True, user_debug_gui=False,
                         vision_attributes=False, **batch_kwargs)
        if act == ActionType.VEL:
            self.SPEED_LIMIT = 0.03 * self.MAX_SPEED_KMH * (1000 / 3600)

    def _actionCode(self):
        return self.ACT_TYPE.value

    def _actionSize(self):
        if self.ACT_TYPE in [ActionType.RPM, ActionType.VEL]:
            return 4
        if self.ACT_TYPE == ActionType.PID:
            return 3
        if self.ACT_TYPE in [ActionType.ONE_D_RPM, ActionType.ONE_D_PID]:
            return 1
        raise ValueError("[ERROR] in BaseRLAviary._actionSpace()")

    def _actionSpace(self):
        size = self._actionSize()
        return Box(low=-np.ones((self.NUM_DRONES, size)), high=np.ones((self.NUM_DRONES, size)), dtype=np.float32)

    def _observationSpace(self):
        lo, hi = -np.inf, np.inf
        n, size = self.NUM_DRONES, self._actionSize()
        low = np.hstack([np.array([[lo, lo, 0, lo, lo, lo, lo, lo, lo, lo, lo, lo] for _ in range(n)]),
                         -np.ones((n, size * self.ACTION_BUFFER_SIZE))])
        high = np.hstack([np.full((n, 12), hi), np.ones((n, size * self.ACTION_BUFFER_SIZE))])
        return Box(low=low, high=high, dtype=np.float32)
