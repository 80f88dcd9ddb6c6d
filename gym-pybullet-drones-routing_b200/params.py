"""Drone constants: URDF parsing and the derived quantities of ``BaseAviary.__init__``.

Mirrors reference ``envs/BaseAviary.py:982-1014`` (``_parseURDFParameters``: same
17-tuple, same order), ``BaseAviary.py:74-83,117-128`` (derived constants, computed
in float64 on the host exactly as there and handed to the kernels as-is — nothing is
recomputed on the device) and ``control/BaseControl.py:181-216``.

Unlike the reference, elements are looked up by tag/attribute name rather than by
positional index, so both this package's minimal ``assets/*.urdf`` and the reference's
full URDFs parse to the same values (tests/test_params.py pins this).
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as etxml
from dataclasses import dataclass

import numpy as np

from .utils.enums import DroneModel

ASSETS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def urdf_path(drone_model: DroneModel) -> str:
    return os.path.join(ASSETS_DIR, drone_model.value + ".urdf")


def parse_urdf_parameters(path: str):
    """Returns ``(M, L, THRUST2WEIGHT_RATIO, J, J_INV, KF, KM, COLLISION_H, COLLISION_R,
    COLLISION_Z_OFFSET, MAX_SPEED_KMH, GND_EFF_COEFF, PROP_RADIUS, DRAG_COEFF,
    DW_COEFF_1, DW_COEFF_2, DW_COEFF_3)`` — reference ``BaseAviary.py:1013-1014``."""
    root = etxml.parse(path).getroot()
    prop = root.find("properties").attrib
    base = root.find("link")
    inertial = base.find("inertial")
    M = float(inertial.find("mass").attrib["value"])
    L = float(prop["arm"])
    T2W = float(prop["thrust2weight"])
    ine = inertial.find("inertia").attrib
    J = np.diag([float(ine["ixx"]), float(ine["iyy"]), float(ine["izz"])])
    J_INV = np.linalg.inv(J)
    KF = float(prop["kf"])
    KM = float(prop["km"])
    col = base.find("collision")
    cyl = col.find("geometry").find("cylinder").attrib
    COLLISION_H = float(cyl["length"])
    COLLISION_R = float(cyl["radius"])
    COLLISION_Z_OFFSET = [float(s) for s in col.find("origin").attrib["xyz"].split()][2]
    MAX_SPEED_KMH = float(prop["max_speed_kmh"])
    GND_EFF_COEFF = float(prop["gnd_eff_coeff"])
    PROP_RADIUS = float(prop["prop_radius"])
    dxy = float(prop["drag_coeff_xy"])
    DRAG_COEFF = np.array([dxy, dxy, float(prop["drag_coeff_z"])])
    return (M, L, T2W, J, J_INV, KF, KM, COLLISION_H, COLLISION_R, COLLISION_Z_OFFSET, MAX_SPEED_KMH,
            GND_EFF_COEFF, PROP_RADIUS, DRAG_COEFF,
            float(prop["dw_coeff_1"]), float(prop["dw_coeff_2"]), float(prop["dw_coeff_3"]))


def parse_rotor_offsets(path: str) -> np.ndarray:
    """(4,3) centre-of-mass offsets of the rotor links ``prop0..3`` in the base frame — what
    ``p.getLinkStates`` resolves for ``_groundEffect`` (reference ``BaseAviary.py:732-739``;
    ``assets/cf2x.urdf:42,54,66,78``)."""
    root = etxml.parse(path).getroot()
    out = []
    for link in root.findall("link")[1:5]:
        org = link.find("inertial").find("origin")
        out.append([float(s) for s in org.attrib.get("xyz", "0 0 0").split()])
    return np.array(out, dtype=np.float64).reshape(4, 3)


@dataclass(frozen=True)
class DroneParams:
    """URDF values + the derived constants of reference ``BaseAviary.py:117-128``."""
    model: DroneModel
    M: float
    L: float
    THRUST2WEIGHT_RATIO: float
    J: np.ndarray
    J_INV: np.ndarray
    KF: float
    KM: float
    COLLISION_H: float
    COLLISION_R: float
    COLLISION_Z_OFFSET: float
    MAX_SPEED_KMH: float
    GND_EFF_COEFF: float
    PROP_RADIUS: float
    DRAG_COEFF: np.ndarray
    DW_COEFF_1: float
    DW_COEFF_2: float
    DW_COEFF_3: float
    ROTOR_XYZ: np.ndarray
    G: float
    GRAVITY: float
    HOVER_RPM: float
    MAX_RPM: float
    MAX_THRUST: float
    MAX_XY_TORQUE: float
    MAX_Z_TORQUE: float
    GND_EFF_H_CLIP: float

    def urdf_tuple(self):
        return (self.M, self.L, self.THRUST2WEIGHT_RATIO, self.J, self.J_INV, self.KF, self.KM, self.COLLISION_H,
                self.COLLISION_R, self.COLLISION_Z_OFFSET, self.MAX_SPEED_KMH, self.GND_EFF_COEFF, self.PROP_RADIUS,
                self.DRAG_COEFF, self.DW_COEFF_1, self.DW_COEFF_2, self.DW_COEFF_3)


def load_drone_params(drone_model: DroneModel, g: float = 9.8, path: str | None = None) -> DroneParams:
    path = path or urdf_path(drone_model)
    (M, L, T2W, J, J_INV, KF, KM, CH, CR, CZ, VMAX, GND, PRAD, DRAG, DW1, DW2, DW3) = parse_urdf_parameters(path)
    GRAVITY = g * M                                              # BaseAviary.py:117
    HOVER_RPM = np.sqrt(GRAVITY / (4 * KF))                      # :118
    MAX_RPM = np.sqrt((T2W * GRAVITY) / (4 * KF))                # :119
    MAX_THRUST = (4 * KF * MAX_RPM ** 2)                         # :120
    if drone_model == DroneModel.CF2P:                           # :121-126
        MAX_XY_TORQUE = (L * KF * MAX_RPM ** 2)
    else:
        MAX_XY_TORQUE = (2 * L * KF * MAX_RPM ** 2) / np.sqrt(2)
    MAX_Z_TORQUE = (2 * KM * MAX_RPM ** 2)                       # :127
    GND_EFF_H_CLIP = 0.25 * PRAD * np.sqrt((15 * MAX_RPM ** 2 * KF * GND) / MAX_THRUST)   # :128
    return DroneParams(drone_model, M, L, T2W, J, J_INV, KF, KM, CH, CR, CZ, VMAX, GND, PRAD, DRAG, DW1, DW2, DW3,
                       parse_rotor_offsets(path), g, float(GRAVITY), float(HOVER_RPM), float(MAX_RPM),
                       float(MAX_THRUST), float(MAX_XY_TORQUE), float(MAX_Z_TORQUE), float(GND_EFF_H_CLIP))


@dataclass
class PIDParams:
    """Gains, PWM map and mixer of reference ``control/DSLPIDControl.py:37-60`` and the
    controller's own ``GRAVITY``/``KF`` (``control/BaseControl.py:35-39``)."""
    P_COEFF_FOR: np.ndarray
    I_COEFF_FOR: np.ndarray
    D_COEFF_FOR: np.ndarray
    P_COEFF_TOR: np.ndarray
    I_COEFF_TOR: np.ndarray
    D_COEFF_TOR: np.ndarray
    PWM2RPM_SCALE: float
    PWM2RPM_CONST: float
    MIN_PWM: float
    MAX_PWM: float
    MIXER_MATRIX: np.ndarray
    GRAVITY: float
    KF: float


def default_pid_params(drone_model: DroneModel, g: float = 9.8) -> PIDParams:
    if drone_model not in (DroneModel.CF2X, DroneModel.CF2P):
        raise ValueError("DSLPIDControl requires DroneModel.CF2X or DroneModel.CF2P")   # DSLPIDControl.py:34-36
    d = load_drone_params(drone_model, g)
    if drone_model == DroneModel.CF2X:
        mixer = np.array([[-.5, -.5, -1], [-.5, .5, 1], [.5, .5, -1], [.5, -.5, 1]], dtype=np.float64)
    else:
        mixer = np.array([[0, -1, -1], [+1, 0, 1], [0, 1, -1], [-1, 0, 1]], dtype=np.float64)
    return PIDParams(np.array([.4, .4, 1.25]), np.array([.05, .05, .05]), np.array([.2, .2, .5]),
                     np.array([70000., 70000., 60000.]), np.array([.0, .0, 500.]), np.array([20000., 20000., 12000.]),
                     0.2685, 4070.3, 20000., 65535., mixer, g * d.M, d.KF)
