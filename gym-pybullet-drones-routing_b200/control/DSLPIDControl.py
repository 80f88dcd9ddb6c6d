"""Batched ``DSLPIDControl`` (reference ``control/DSLPIDControl.py``): one object holds the state of ``n``
independent controllers on the device; ``computeControl`` is one launch of ``gpd_pid_compute``."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..params import PIDParams, default_pid_params
from ..utils.enums import DroneModel
from .BaseControl import BaseControl


class DSLPIDControl(BaseControl):
    def __init__(self, drone_model: DroneModel, g: float = 9.8, num: int = 1, device: int = 0, precision: str = "f64"):
        super().__init__(drone_model=drone_model, g=g)
        if self.DRONE_MODEL != DroneModel.CF2X and self.DRONE_MODEL != DroneModel.CF2P:
            raise ValueError("[ERROR] in DSLPIDControl.__init__(), DSLPIDControl requires DroneModel.CF2X or DroneModel.CF2P")
        d = default_pid_params(drone_model, g)                      # DSLPIDControl.py:37-60
        self.P_COEFF_FOR, self.I_COEFF_FOR, self.D_COEFF_FOR = d.P_COEFF_FOR, d.I_COEFF_FOR, d.D_COEFF_FOR
        self.P_COEFF_TOR, self.I_COEFF_TOR, self.D_COEFF_TOR = d.P_COEFF_TOR, d.I_COEFF_TOR, d.D_COEFF_TOR
        self.PWM2RPM_SCALE, self.PWM2RPM_CONST = d.PWM2RPM_SCALE, d.PWM2RPM_CONST
        self.MIN_PWM, self.MAX_PWM = d.MIN_PWM, d.MAX_PWM
        self.MIXER_MATRIX = d.MIXER_MATRIX
        self.num = int(num)
        self.precision = precision
        self.real = torch.float64 if precision == "f64" else torch.float32
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self._lib = _lib.load()
        self.state = None
        self.reset()

    _PARAM_NAMES = frozenset(("P_COEFF_FOR", "I_COEFF_FOR", "D_COEFF_FOR", "P_COEFF_TOR", "I_COEFF_TOR", "D_COEFF_TOR",
                              "PWM2RPM_SCALE", "PWM2RPM_CONST", "MIN_PWM", "MAX_PWM", "MIXER_MATRIX", "GRAVITY", "KF"))

    def __setattr__(self, name, value):
        # the C parameter block is rebuilt only after a gain / mixer attribute was assigned (setPIDCoefficients or directly)
        if name in DSLPIDControl._PARAM_NAMES:
            object.__setattr__(self, "_pc", None)
        object.__setattr__(self, name, value)

    def reset(self):
        """DSLPIDControl.py:65-78: integral errors and last rpy to zero."""
        super().reset()
        if getattr(self, "device", None) is not None:
            self.state = torch.zeros((self.num, 9), dtype=self.real, device=self.device)

    # views with the reference's attribute names
    integral_pos_e = property(lambda self: self.state[:, 0:3])
    integral_rpy_e = property(lambda self: self.state[:, 3:6])
    last_rpy = property(lambda self: self.state[:, 6:9])

    def _params(self) -> PIDParams:
        return PIDParams(np.asarray(self.P_COEFF_FOR, float), np.asarray(self.I_COEFF_FOR, float),
                         np.asarray(self.D_COEFF_FOR, float), np.asarray(self.P_COEFF_TOR, float),
                         np.asarray(self.I_COEFF_TOR, float), np.asarray(self.D_COEFF_TOR, float),
                         self.PWM2RPM_SCALE, self.PWM2RPM_CONST, self.MIN_PWM, self.MAX_PWM,
                         np.asarray(self.MIXER_MATRIX, float), self.GRAVITY, self.KF)

    def _t(self, x, cols):
        if x is None:
            return None
        t = torch.as_tensor(x, dtype=self.real, device=self.device).reshape(-1, cols)
        if t.shape[0] == 1 and self.num > 1:
            t = t.expand(self.num, cols)
        if t.shape[0] != self.num:
            raise ValueError(f"expected {self.num} rows of {cols}, got {tuple(t.shape)}")
        return t.contiguous()

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos, target_rpy=None,
                       target_vel=None, target_rpy_rates=None):
        """(rpm (n,4), pos_e (n,3), yaw_e (n,)) — DSLPIDControl.py:82-145.  ``cur_ang_vel`` is unused there too."""
        self.control_counter += 1
        n = self.num
        cp, cq, cv, tp = self._t(cur_pos, 3), self._t(cur_quat, 4), self._t(cur_vel, 3), self._t(target_pos, 3)
        tr, tv, trr = self._t(target_rpy, 3), self._t(target_vel, 3), self._t(target_rpy_rates, 3)
        out = torch.empty((8, n), dtype=self.real, device=self.device)      # one block: fresh results every call, like the reference
        rpm, pos_e, yaw_e = out[0:4].view(n, 4), out[4:7].view(n, 3), out[7]
        pc = getattr(self, "_pc", None)
        if pc is None:
            pc = _lib.pid_params_c(self._params())
            object.__setattr__(self, "_pc", pc)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.gpd_pid_compute(self.device_index, _lib.GPD_F64 if self.precision == "f64" else _lib.GPD_F32,
                                             C.byref(pc), n, float(control_timestep), p(cp), p(cq), p(cv), p(tp), p(tr),
                                             p(tv), p(trr), p(self.state), p(rpm), p(pos_e), p(yaw_e), st))
        self._keepalive = (cp, cq, cv, tp, tr, tv, trr)
        return rpm, pos_e, yaw_e

    def _one23DInterface(self, thrust):
        """1, 2 or 4 thrust inputs -> 4 motor PWMs (reference DSLPIDControl.py:263-287); ``thrust``: (..., DIM)."""
        t = torch.as_tensor(thrust, dtype=self.real, device=self.device)
        DIM = t.shape[-1]
        pwm = torch.clamp((torch.sqrt(t / (self.KF * (4 / DIM))) - self.PWM2RPM_CONST) / self.PWM2RPM_SCALE,
                          self.MIN_PWM, self.MAX_PWM)
        if DIM in (1, 4):
            return torch.repeat_interleave(pwm, 4 // DIM, dim=-1)
        if DIM == 2:
            return torch.cat([pwm, torch.flip(pwm, dims=(-1,))], dim=-1)
        raise ValueError("[ERROR] in DSLPIDControl._one23DInterface()")
