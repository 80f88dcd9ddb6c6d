"""Batched ``BaseControl`` (reference ``control/BaseControl.py``): URDF-derived ``GRAVITY``/``KF``/``KM``
(``:35-39``), ``computeControlFromState`` (``:55-93``), ``setPIDCoefficients`` (``:138-177``)."""
import numpy as np

from ..params import load_drone_params
from ..utils.enums import DroneModel


class BaseControl(object):
    def __init__(self, drone_model: DroneModel, g: float = 9.8):
        self.DRONE_MODEL = drone_model
        p = load_drone_params(drone_model, g)
        self.GRAVITY = g * p.M
        self.KF = p.KF
        self.KM = p.KM
        self.reset()

    def reset(self):
        self.control_counter = 0

    def computeControlFromState(self, control_timestep, state, target_pos, target_rpy=None, target_vel=None,
                                target_rpy_rates=None):
        """``state``: (n, 20) rows as returned by ``BaseAviary._getDroneStateVector`` / the Ctrl observation."""
        return self.computeControl(control_timestep=control_timestep, cur_pos=state[..., 0:3], cur_quat=state[..., 3:7],
                                   cur_vel=state[..., 10:13], cur_ang_vel=state[..., 13:16], target_pos=target_pos,
                                   target_rpy=target_rpy, target_vel=target_vel, target_rpy_rates=target_rpy_rates)

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos, target_rpy=None,
                       target_vel=None, target_rpy_rates=None):
        raise NotImplementedError

    def setPIDCoefficients(self, p_coeff_pos=None, i_coeff_pos=None, d_coeff_pos=None, p_coeff_att=None,
                           i_coeff_att=None, d_coeff_att=None):
        ATTR_LIST = ['P_COEFF_FOR', 'I_COEFF_FOR', 'D_COEFF_FOR', 'P_COEFF_TOR', 'I_COEFF_TOR', 'D_COEFF_TOR']
        if not all(hasattr(self, attr) for attr in ATTR_LIST):
            raise AttributeError("[ERROR] in BaseControl.setPIDCoefficients(), not all PID coefficients exist as "
                                 "attributes in the instantiated control class.")
        for name, val in zip(ATTR_LIST, [p_coeff_pos, i_coeff_pos, d_coeff_pos, p_coeff_att, i_coeff_att, d_coeff_att]):
            if val is not None:
                setattr(self, name, np.asarray(val, dtype=np.float64))
