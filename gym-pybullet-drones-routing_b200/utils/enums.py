"""Enumerations of the batched simulator.

Same member names and string values as the reference's
``gym_pybullet_drones/utils/enums.py:3-48`` so ``DroneModel("cf2x")``,
``Physics("dyn")`` etc. keep working.  ``Physics`` gains the DYN-form composites
(explicit dynamics + the closed-form ground-effect / drag / downwash models of
``BaseAviary.py:715-811``), which have no counterpart in the reference: there the
three models only feed Bullet's solver (``PYB_*``), which is out of scope here.
"""
from enum import Enum


class DroneModel(Enum):
    CF2X = "cf2x"
    CF2P = "cf2p"
    RACE = "racer"


class Physics(Enum):
    PYB = "pyb"
    DYN = "dyn"
    PYB_GND = "pyb_gnd"
    PYB_DRAG = "pyb_drag"
    PYB_DW = "pyb_dw"
    PYB_GND_DRAG_DW = "pyb_gnd_drag_dw"
    # --- new: explicit dynamics with the closed-form force models injected ---
    DYN_GND = "dyn_gnd"
    DYN_DRAG = "dyn_drag"
    DYN_DW = "dyn_dw"
    DYN_GND_DRAG = "dyn_gnd_drag"
    DYN_GND_DRAG_DW = "dyn_gnd_drag_dw"


class ImageType(Enum):
    RGB = 0
    DEP = 1
    SEG = 2
    BW = 3


class ActionType(Enum):
    RPM = "rpm"
    PID = "pid"
    VEL = "vel"
    ONE_D_RPM = "one_d_rpm"
    ONE_D_PID = "one_d_pid"


class ObservationType(Enum):
    KIN = "kin"
    RGB = "rgb"


#: C-ABI bit flags (include/gpd.h GPD_PHY_*) of the force models each DYN-form mode enables
PHYSICS_FLAGS = {
    Physics.DYN: 0,
    Physics.DYN_GND: 1,
    Physics.DYN_DRAG: 2,
    Physics.DYN_DW: 4,
    Physics.DYN_GND_DRAG: 1 | 2,
    Physics.DYN_GND_DRAG_DW: 1 | 2 | 4,
}
