"""ctypes binding of the CPU oracle (oracle/libgpd_oracle.so).

TEST INFRASTRUCTURE — imported only by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgpd_oracle.so")

ACT = {"rpm": 0, "pid": 1, "vel": 2, "one_d_rpm": 3, "one_d_pid": 4, "ctrl_rpm": 5, "ctrl_vel": 6}
ENV = {"ctrl": 0, "hover": 1, "multihover": 2}
MODEL = {"cf2x": 0, "cf2p": 1, "racer": 2}
PHY_GND, PHY_DRAG, PHY_DW = 1, 2, 4


class OrcDrone(C.Structure):
    _fields_ = [("model", C.c_int32), ("_pad", C.c_int32),
                ("M", C.c_double), ("L", C.c_double), ("THRUST2WEIGHT", C.c_double),
                ("J", C.c_double * 3), ("J_INV", C.c_double * 3),
                ("KF", C.c_double), ("KM", C.c_double),
                ("COLLISION_H", C.c_double), ("COLLISION_R", C.c_double), ("COLLISION_Z_OFFSET", C.c_double),
                ("MAX_SPEED_KMH", C.c_double), ("GND_EFF_COEFF", C.c_double), ("PROP_RADIUS", C.c_double),
                ("DRAG_COEFF", C.c_double * 3),
                ("DW_COEFF_1", C.c_double), ("DW_COEFF_2", C.c_double), ("DW_COEFF_3", C.c_double),
                ("G", C.c_double), ("GRAVITY", C.c_double), ("HOVER_RPM", C.c_double), ("MAX_RPM", C.c_double),
                ("MAX_THRUST", C.c_double), ("MAX_XY_TORQUE", C.c_double), ("MAX_Z_TORQUE", C.c_double),
                ("GND_EFF_H_CLIP", C.c_double),
                ("ROTOR_XYZ", (C.c_double * 3) * 4)]


class OrcPid(C.Structure):
    _fields_ = [("P_FOR", C.c_double * 3), ("I_FOR", C.c_double * 3), ("D_FOR", C.c_double * 3),
                ("P_TOR", C.c_double * 3), ("I_TOR", C.c_double * 3), ("D_TOR", C.c_double * 3),
                ("PWM2RPM_SCALE", C.c_double), ("PWM2RPM_CONST", C.c_double),
                ("MIN_PWM", C.c_double), ("MAX_PWM", C.c_double),
                ("MIXER", (C.c_double * 3) * 4),
                ("GRAVITY", C.c_double), ("KF", C.c_double)]


class OrcEnvCfg(C.Structure):
    _fields_ = [("num_drones", C.c_int32), ("substeps", C.c_int32), ("pyb_freq", C.c_int32), ("ctrl_freq", C.c_int32),
                ("env_kind", C.c_int32), ("action_type", C.c_int32), ("physics_flags", C.c_int32),
                ("action_buffer_size", C.c_int32),
                ("episode_len_sec", C.c_double), ("speed_limit", C.c_double),
                ("drone", OrcDrone), ("pid", OrcPid)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "gpd_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "gpd_oracle.h"))):
        subprocess.run(["make", "-C", _HERE, "-B", "libgpd_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.orc_matrix_from_quaternion.argtypes = [dp, dp]
        L.orc_euler_from_quaternion.argtypes = [dp, dp]
        L.orc_quaternion_from_euler.argtypes = [dp, dp]
        L.orc_integrate_q.argtypes = [dp, dp, C.c_double, dp]
        L.orc_dynamics.argtypes = [C.POINTER(OrcDrone), C.c_double, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_ground_effect.argtypes = [C.POINTER(OrcDrone), dp, dp, dp, dp, dp]
        L.orc_ground_effect.restype = C.c_int
        L.orc_drag.argtypes = [C.POINTER(OrcDrone), dp, dp, dp, dp]
        L.orc_downwash.argtypes = [C.POINTER(OrcDrone), C.c_int, dp, C.c_int]
        L.orc_downwash.restype = C.c_double
        L.orc_pid_compute.argtypes = [C.POINTER(OrcPid), C.c_double, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_calculate_next_step.argtypes = [dp, dp, C.c_double, dp]
        L.orc_step.argtypes = [C.POINTER(OrcEnvCfg), C.c_int64, dp, dp, dp, C.POINTER(C.c_float),
                               C.POINTER(C.c_int32), C.c_void_p, dp, C.c_void_p, dp,
                               C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int]
        L.orc_reset.argtypes = [C.POINTER(OrcEnvCfg), C.c_int64, C.POINTER(C.c_uint8), dp, dp, dp, dp,
                                C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_void_p]
        L.orc_action_width.argtypes = [C.c_int]
        L.orc_action_width.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if n is not None:
        assert a.size == n, (a.shape, n)
    return a


def make_drone(p) -> OrcDrone:
    """``p``: any object with the attribute names of the reference's BaseAviary constants."""
    d = OrcDrone()
    model = getattr(p, "model", None)
    d.model = MODEL[model.value if hasattr(model, "value") else model]
    d.M, d.L, d.THRUST2WEIGHT = p.M, p.L, p.THRUST2WEIGHT_RATIO
    J, JI = np.asarray(p.J), np.asarray(p.J_INV)
    for k in range(3):
        d.J[k] = J[k, k]
        d.J_INV[k] = JI[k, k]
        d.DRAG_COEFF[k] = float(np.asarray(p.DRAG_COEFF)[k])
    d.KF, d.KM = p.KF, p.KM
    d.COLLISION_H, d.COLLISION_R, d.COLLISION_Z_OFFSET = p.COLLISION_H, p.COLLISION_R, p.COLLISION_Z_OFFSET
    d.MAX_SPEED_KMH, d.GND_EFF_COEFF, d.PROP_RADIUS = p.MAX_SPEED_KMH, p.GND_EFF_COEFF, p.PROP_RADIUS
    d.DW_COEFF_1, d.DW_COEFF_2, d.DW_COEFF_3 = p.DW_COEFF_1, p.DW_COEFF_2, p.DW_COEFF_3
    d.G, d.GRAVITY, d.HOVER_RPM, d.MAX_RPM = p.G, p.GRAVITY, p.HOVER_RPM, p.MAX_RPM
    d.MAX_THRUST, d.MAX_XY_TORQUE, d.MAX_Z_TORQUE, d.GND_EFF_H_CLIP = (
        p.MAX_THRUST, p.MAX_XY_TORQUE, p.MAX_Z_TORQUE, p.GND_EFF_H_CLIP)
    rx = np.asarray(p.ROTOR_XYZ, dtype=np.float64).reshape(4, 3)
    for i in range(4):
        for k in range(3):
            d.ROTOR_XYZ[i][k] = rx[i, k]
    return d


def make_pid(c) -> OrcPid:
    """``c``: any object with the attribute names of the reference's DSLPIDControl."""
    o = OrcPid()
    for k in range(3):
        o.P_FOR[k], o.I_FOR[k], o.D_FOR[k] = c.P_COEFF_FOR[k], c.I_COEFF_FOR[k], c.D_COEFF_FOR[k]
        o.P_TOR[k], o.I_TOR[k], o.D_TOR[k] = c.P_COEFF_TOR[k], c.I_COEFF_TOR[k], c.D_COEFF_TOR[k]
    o.PWM2RPM_SCALE, o.PWM2RPM_CONST = c.PWM2RPM_SCALE, c.PWM2RPM_CONST
    o.MIN_PWM, o.MAX_PWM = c.MIN_PWM, c.MAX_PWM
    mx = np.asarray(c.MIXER_MATRIX, dtype=np.float64)
    for i in range(4):
        for k in range(3):
            o.MIXER[i][k] = mx[i, k]
    o.GRAVITY, o.KF = c.GRAVITY, c.KF
    return o


def matrix_from_quaternion(q):
    q = _f64(q, 4)
    m = np.empty(9)
    lib().orc_matrix_from_quaternion(_dp(q), _dp(m))
    return m.reshape(3, 3)


def euler_from_quaternion(q):
    q = _f64(q, 4)
    e = np.empty(3)
    lib().orc_euler_from_quaternion(_dp(q), _dp(e))
    return e


def quaternion_from_euler(rpy):
    r = _f64(rpy, 3)
    q = np.empty(4)
    lib().orc_quaternion_from_euler(_dp(r), _dp(q))
    return q


def ground_effect(drone: OrcDrone, rpm, pos, quat, rpy):
    out = np.empty(4)
    ok = lib().orc_ground_effect(C.byref(drone), _dp(_f64(rpm, 4)), _dp(_f64(pos, 3)), _dp(_f64(quat, 4)),
                                 _dp(_f64(rpy, 3)), _dp(out))
    return out, bool(ok)


def drag(drone: OrcDrone, rpm_prev, quat, vel):
    out = np.empty(3)
    lib().orc_drag(C.byref(drone), _dp(_f64(rpm_prev, 4)), _dp(_f64(quat, 4)), _dp(_f64(vel, 3)), _dp(out))
    return out


def downwash(drone: OrcDrone, pos_all, i):
    pa = _f64(pos_all)
    return lib().orc_downwash(C.byref(drone), pa.shape[0], _dp(pa), int(i))


def pid_compute(pid: OrcPid, dt, cur_pos, cur_quat, cur_vel, target_pos, target_rpy=None, target_vel=None,
                target_rpy_rates=None, pid_state=None):
    """One DSLPIDControl.computeControl call; ``pid_state`` (9,) is updated in place."""
    z = np.zeros(3)
    st = pid_state if pid_state is not None else np.zeros(9)
    assert st.dtype == np.float64 and st.size == 9 and st.flags.c_contiguous
    rpm, pos_e, yaw = np.empty(4), np.empty(3), C.c_double(0)
    lib().orc_pid_compute(C.byref(pid), float(dt), _dp(_f64(cur_pos, 3)), _dp(_f64(cur_quat, 4)),
                          _dp(_f64(cur_vel, 3)), _dp(_f64(target_pos, 3)),
                          _dp(_f64(z if target_rpy is None else target_rpy, 3)),
                          _dp(_f64(z if target_vel is None else target_vel, 3)),
                          _dp(_f64(z if target_rpy_rates is None else target_rpy_rates, 3)),
                          _dp(st), _dp(rpm), _dp(pos_e), C.byref(yaw))
    return rpm, pos_e, yaw.value


class OracleSim:
    """Batched CPU env (E independent envs × N drones) driven by ``orc_step``/``orc_reset``.

    Mirrors the reference's BaseAviary/BaseRLAviary/Hover/MultiHover/Ctrl step contract on
    Physics.DYN (+ DYN-form force models) in FP64.
    """

    def __init__(self, drone_params, num_envs, num_drones=1, env_kind="hover", action_type="rpm",
                 pyb_freq=240, ctrl_freq=30, physics_flags=0, pid_params=None, init_xyz=None, init_rpy=None,
                 target_pos=None, episode_len_sec=8.0):
        self.E, self.N = int(num_envs), int(num_drones)
        cfg = OrcEnvCfg()
        cfg.num_drones = self.N
        if pyb_freq % ctrl_freq != 0:
            raise ValueError("pyb_freq is not divisible by ctrl_freq")
        cfg.substeps = pyb_freq // ctrl_freq
        cfg.pyb_freq, cfg.ctrl_freq = pyb_freq, ctrl_freq
        cfg.env_kind = ENV[env_kind]
        cfg.action_type = ACT[action_type]
        cfg.physics_flags = physics_flags
        self.is_ctrl = env_kind == "ctrl"
        cfg.action_buffer_size = 0 if self.is_ctrl else ctrl_freq // 2
        cfg.episode_len_sec = episode_len_sec
        cfg.speed_limit = 0.03 * drone_params.MAX_SPEED_KMH * (1000 / 3600)
        cfg.drone = make_drone(drone_params)
        if pid_params is not None:
            cfg.pid = make_pid(pid_params)
        self.cfg = cfg
        self.A = lib().orc_action_width(cfg.action_type)
        self.B = cfg.action_buffer_size
        self.W = 20 if self.is_ctrl else 12 + self.A * self.B
        E, N = self.E, self.N
        p = drone_params
        if init_xyz is None:                                     # BaseAviary.py:194-197
            one = np.stack([np.array([x * 4 * p.L for x in range(N)]), np.array([y * 4 * p.L for y in range(N)]),
                            np.ones(N) * (p.COLLISION_H / 2 - p.COLLISION_Z_OFFSET + .1)], axis=1)
            init_xyz = one
        if init_rpy is None:
            init_rpy = np.zeros((N, 3))
        self.init_xyz = np.ascontiguousarray(np.broadcast_to(np.asarray(init_xyz, np.float64), (E, N, 3)))
        self.init_rpy = np.ascontiguousarray(np.broadcast_to(np.asarray(init_rpy, np.float64), (E, N, 3)))
        if target_pos is None:
            if env_kind == "hover":
                target_pos = np.array([[0., 0., 1.]])                                      # HoverAviary.py:51
            else:
                target_pos = self.init_xyz[0] + np.array([[0, 0, 1 / (i + 1)] for i in range(N)])   # MultiHoverAviary.py:71
        self.target_pos = _f64(target_pos, N * 3).reshape(N, 3)
        self.state20 = np.zeros((E, N, 20))
        self.rpy_rates = np.zeros((E, N, 3))
        self.pid_state = np.zeros((E, N, 9))
        self.ring = np.zeros((E, N, max(self.B, 1), self.A), dtype=np.float32)
        self.step_counter = np.zeros(E, dtype=np.int32)
        self.obs = np.zeros((E, N, self.W), dtype=np.float64 if self.is_ctrl else np.float32)
        self.reward = np.zeros(E)
        self.terminated = np.zeros(E, dtype=np.uint8)
        self.truncated = np.zeros(E, dtype=np.uint8)
        self.reset()

    def reset(self, mask=None):
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
        lib().orc_reset(C.byref(self.cfg), self.E, None if m is None else m.ctypes.data_as(C.POINTER(C.c_uint8)),
                        _dp(self.init_xyz), _dp(self.init_rpy), _dp(self.state20), _dp(self.rpy_rates),
                        self.ring.ctypes.data_as(C.POINTER(C.c_float)),
                        self.step_counter.ctypes.data_as(C.POINTER(C.c_int32)), self.obs.ctypes.data_as(C.c_void_p))
        return self.obs

    def step(self, actions, nthreads=1):
        dt = np.float64 if self.is_ctrl else np.float32
        a = np.ascontiguousarray(np.asarray(actions, dtype=dt).reshape(self.E, self.N, self.A))
        lib().orc_step(C.byref(self.cfg), self.E, _dp(self.state20), _dp(self.rpy_rates), _dp(self.pid_state),
                       self.ring.ctypes.data_as(C.POINTER(C.c_float)),
                       self.step_counter.ctypes.data_as(C.POINTER(C.c_int32)),
                       a.ctypes.data_as(C.c_void_p), _dp(self.target_pos), self.obs.ctypes.data_as(C.c_void_p),
                       _dp(self.reward), self.terminated.ctypes.data_as(C.POINTER(C.c_uint8)),
                       self.truncated.ctypes.data_as(C.POINTER(C.c_uint8)), int(nthreads))
        return self.obs, self.reward, self.terminated, self.truncated


def max_threads() -> int:
    return lib().orc_max_threads()


def adjacency_matrix(pos, neighbourhood_radius):
    """BaseAviary._getAdjacencyMatrix (BaseAviary.py:658-675) restated for one env: identity, plus 1 where
    ``np.linalg.norm(pos[i] - pos[j]) < NEIGHBOURHOOD_RADIUS`` for i < j (mirrored).  ``pos``: (N, 3) float64."""
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[0]
    adj = np.identity(n)
    for i in range(n - 1):
        for j in range(n - i - 1):
            if np.linalg.norm(pos[i, :] - pos[j + i + 1, :]) < neighbourhood_radius:
                adj[i, j + i + 1] = adj[j + i + 1, i] = 1
    return adj
