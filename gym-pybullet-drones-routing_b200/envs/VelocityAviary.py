"""Batched ``VelocityAviary`` (reference ``envs/VelocityAviary.py``): each drone follows a commanded velocity vector
``(vx, vy, vz, speed fraction)`` through its in-env ``DSLPIDControl`` (``:129-170``); observation = 20-float state
(``:117-127``), dummy reward/flags (``:174-228``).  ``SPEED_LIMIT = 0.03 * MAX_SPEED_KMH / 3.6`` (``:78``)."""
import numpy as np

from ..spaces import Box
from ..utils.enums import DroneModel, Physics
from .CtrlAviary import CtrlAviary


class VelocityAviary(CtrlAviary):
    ENV_KIND = "ctrl"

    def __init__(self, drone_model: DroneModel = DroneModel.CF2X, num_drones: int = 1,
                 neighbourhood_radius: float = np.inf, initial_xyzs=None, initial_rpys=None,
                 physics: Physics = Physics.DYN, pyb_freq: int = 240, ctrl_freq: int = 240, gui=False, record=False,
                 obstacles=False, user_debug_gui=True, output_folder='results', **batch_kwargs):
        if drone_model not in [DroneModel.CF2X, DroneModel.CF2P]:
            raise ValueError("VelocityAviary needs DSLPIDControl, which exists for DroneModel.CF2X / CF2P only")
        super().__init__(drone_model=drone_model, num_drones=num_drones, neighbourhood_radius=neighbourhood_radius,
                         initial_xyzs=initial_xyzs, initial_rpys=initial_rpys, physics=physics, pyb_freq=pyb_freq,
                         ctrl_freq=ctrl_freq, gui=gui, record=record, obstacles=obstacles,
                         user_debug_gui=user_debug_gui, output_folder=output_folder, **batch_kwargs)
        self.SPEED_LIMIT = 0.03 * self.MAX_SPEED_KMH * (1000 / 3600)

    def _actionCode(self):
        return "ctrl_vel"

    def _actionSpace(self):
        lo = np.array([[-1, -1, -1, 0] for _ in range(self.NUM_DRONES)])
        hi = np.array([[1, 1, 1, 1] for _ in range(self.NUM_DRONES)])
        return Box(low=lo, high=hi, dtype=np.float32)
