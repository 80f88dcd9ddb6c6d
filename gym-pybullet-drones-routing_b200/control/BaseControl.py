"""Batched ``BaseControl`` (reference ``control/BaseControl.py``): URDF-derived ``GRAVITY``/``KF``/``KM``
(``:35-39``), ``computeControlFromState`` (``:55-93``), ``setPIDCoefficients`` (``:138-177``)."""
import numpy as np

import xml.etree.ElementTree as etxml

from ..params import load_drone_params, urdf_path
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

    def _getURDFParameter(self, parameter_name: str):
        """Reads one parameter of the controller's drone model from its URDF (reference BaseControl.py:181-216)."""
        root = etxml.parse(urdf_path(self.DRONE_MODEL)).getroot()
        base = root.find("link")
        if parameter_name == 'm':
            return float(base.find("inertial").find("mass").attrib['value'])
        if parameter_name in ['ixx', 'iyy', 'izz']:
            return float(base.find("inertial").find("inertia").attrib[parameter_name])
        if parameter_name in ['arm', 'thrust2weight', 'kf', 'km', 'max_speed_kmh', 'gnd_eff_coeff', 'prop_radius',
                              'drag_coeff_xy', 'drag_coeff_z', 'dw_coeff_1', 'dw_coeff_2', 'dw_coeff_3']:
            return float(root.find("properties").attrib[parameter_name])
        if parameter_name in ['length', 'radius']:
            return float(base.find("collision").find("geometry").find("cylinder").attrib[parameter_name])
        if parameter_name == 'collision_z_offset':
            return [float(v) for v in base.find("collision").find("origin").attrib['xyz'].split()][2]
        return None
