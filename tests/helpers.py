"""Shared helpers of the parity tests: golden loading, the error metric of SURVEY §8d, and
factories that build the CPU oracle (oracle/) for a golden case."""
import json
import os

import numpy as np

import gpd_b200  # noqa: F401
from gpd_b200.params import default_pid_params, load_drone_params
from gpd_b200.utils.enums import DroneModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

S_POS, S_QUAT, S_RPY, S_VEL, S_ANGV, S_RPM, S_RATES = (slice(0, 3), slice(3, 7), slice(7, 10), slice(10, 13),
                                                       slice(13, 16), slice(16, 20), slice(20, 23))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def constants():
    return json.load(open(os.path.join(GOLDEN, "constants.json")))


def traj_cases():
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith("traj_") or
                  (f.startswith("composite_") and "kat" not in f))


def rel_err(x, ref, floor=1e-3):
    """max over rows of ||x - ref||_2 / max(||ref||_2, floor)   (SURVEY §8d parity protocol)."""
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    num = np.linalg.norm(x - ref, axis=-1)
    den = np.maximum(np.linalg.norm(ref, axis=-1), floor)
    return float(np.max(num / den)) if num.size else 0.0


def quat_err(q, ref):
    """quaternions compared up to sign (SURVEY §8c residual uncertainty)."""
    q, ref = np.asarray(q), np.asarray(ref)
    s = np.sign(np.sum(q * ref, axis=-1, keepdims=True))
    s[s == 0] = 1
    return rel_err(q * s, ref)


def angle_err(a, ref):
    d = np.asarray(a) - np.asarray(ref)
    d = (d + np.pi) % (2 * np.pi) - np.pi
    return float(np.max(np.abs(d))) if d.size else 0.0


def case_setup(g):
    """kwargs describing the env of a golden trajectory file."""
    model = DroneModel(str(g["model"]))
    env = {"HoverAviary": "hover", "MultiHoverAviary": "multihover", "CtrlAviary": "ctrl", "VelocityAviary": "ctrl"}[str(g["env"])]
    kw = dict(model=model, env_kind=env, action_type=str(g["act_type"]), num_drones=int(g["num_drones"]),
              pyb_freq=int(g["pyb_freq"]), ctrl_freq=int(g["ctrl_freq"]),
              physics_flags=int(g["flags"]) if "flags" in g.files else 0,
              init_xyz=g["init_xyz"] if "init_xyz" in g.files else None,
              init_rpy=g["init_rpy"] if "init_rpy" in g.files else None)
    return kw


def default_targets(kw):
    """TARGET_POS of HoverAviary.py:51 / MultiHoverAviary.py:71 from the DEFAULT initial poses."""
    dp = load_drone_params(kw["model"])
    n = kw["num_drones"]
    if kw["env_kind"] == "hover":
        return np.array([[0., 0., 1.]])
    if kw["env_kind"] == "multihover":
        init = np.stack([np.array([x * 4 * dp.L for x in range(n)]), np.array([y * 4 * dp.L for y in range(n)]),
                         np.ones(n) * (dp.COLLISION_H / 2 - dp.COLLISION_Z_OFFSET + .1)], axis=1)
        return init + np.array([[0, 0, 1 / (i + 1)] for i in range(n)])
    return None


def make_oracle(kw, num_envs=1):
    from oracle import oracle as orc
    dp = load_drone_params(kw["model"])
    pid = default_pid_params(DroneModel.CF2X) if kw["action_type"] in ("pid", "vel", "one_d_pid", "ctrl_vel") else None
    return orc.OracleSim(dp, num_envs, num_drones=kw["num_drones"], env_kind=kw["env_kind"],
                         action_type=kw["action_type"], pyb_freq=kw["pyb_freq"], ctrl_freq=kw["ctrl_freq"],
                         physics_flags=kw["physics_flags"], pid_params=pid, init_xyz=kw["init_xyz"],
                         init_rpy=kw["init_rpy"], target_pos=default_targets(kw))
