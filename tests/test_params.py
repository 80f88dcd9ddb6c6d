"""Constants layer vs the reference's _parseURDFParameters / derived constants (golden constants.json)."""
import numpy as np
import pytest

from helpers import constants
from gpd_b200.params import default_pid_params, load_drone_params
from gpd_b200.utils.enums import ActionType, DroneModel, ObservationType, Physics


@pytest.mark.parametrize("model", list(DroneModel))
def test_constants_match_reference(model):
    ref = constants()["models"][model.value]
    p = load_drone_params(model)
    for k in ["M", "L", "THRUST2WEIGHT_RATIO", "KF", "KM", "COLLISION_H", "COLLISION_R", "COLLISION_Z_OFFSET",
              "MAX_SPEED_KMH", "GND_EFF_COEFF", "PROP_RADIUS", "DW_COEFF_1", "DW_COEFF_2", "DW_COEFF_3", "G", "GRAVITY",
              "HOVER_RPM", "MAX_RPM", "MAX_THRUST", "MAX_XY_TORQUE", "MAX_Z_TORQUE", "GND_EFF_H_CLIP"]:
        assert getattr(p, k) == ref[k], k                      # bit-identical float64
    assert np.array_equal(p.J, np.array(ref["J"])) and np.array_equal(p.J_INV, np.array(ref["J_INV"]))
    assert np.array_equal(p.DRAG_COEFF, np.array(ref["DRAG_COEFF"]))
    assert len(p.urdf_tuple()) == 17


@pytest.mark.parametrize("model", [DroneModel.CF2X, DroneModel.CF2P])
def test_pid_constants_match_reference(model):
    ref = constants()["models"][model.value]["pid"]
    c = default_pid_params(model)
    for k, v in ref.items():
        assert np.array_equal(np.asarray(getattr(c, k), np.float64), np.asarray(v, np.float64)), k


def test_pid_rejects_racer():
    with pytest.raises(ValueError):
        default_pid_params(DroneModel.RACE)


def test_enum_values():
    assert DroneModel("cf2x") is DroneModel.CF2X and DroneModel("racer") is DroneModel.RACE
    assert Physics("dyn") is Physics.DYN and Physics("pyb_gnd_drag_dw") is Physics.PYB_GND_DRAG_DW
    assert [a.value for a in ActionType] == ["rpm", "pid", "vel", "one_d_rpm", "one_d_pid"]
    assert [o.value for o in ObservationType] == ["kin", "rgb"]


def test_base_control_urdf_parameter_lookup():
    from gpd_b200.control.BaseControl import BaseControl
    c = BaseControl(DroneModel.CF2X)
    ref = constants()["models"]["cf2x"]
    assert c._getURDFParameter('m') == ref["M"] and c._getURDFParameter('kf') == ref["KF"]
    assert c._getURDFParameter('ixx') == ref["J"][0][0] and c._getURDFParameter('arm') == ref["L"]
    assert c._getURDFParameter('length') == ref["COLLISION_H"] and c._getURDFParameter('collision_z_offset') == 0.0
    assert c.GRAVITY == ref["GRAVITY"] and c.KM == ref["KM"]
