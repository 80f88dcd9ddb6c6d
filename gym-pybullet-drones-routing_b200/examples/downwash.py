"""Batched counterpart of the reference's ``examples/downwash.py``: two drones crossing in the X-Z plane, the upper
one's wake pushing the lower one down — here on ``Physics.DYN_DW`` (the reference uses ``Physics.PYB_DW``).

    python -m gpd_b200.examples.downwash --num_envs 1024
"""
import argparse

import numpy as np
import torch

from ..control.DSLPIDControl import DSLPIDControl
from ..envs.CtrlAviary import CtrlAviary
from ..utils.enums import DroneModel, Physics


def run(drone=DroneModel("cf2p"), num_envs=256, simulation_freq_hz=240, control_freq_hz=48, duration_sec=12, precision="f64",
        physics=Physics.DYN_DW):
    INIT_XYZS = np.array([[.5, 0, 1], [-.5, 0, .5]])
    env = CtrlAviary(drone_model=drone, num_drones=2, initial_xyzs=INIT_XYZS, physics=physics, pyb_freq=simulation_freq_hz,
                     ctrl_freq=control_freq_hz, num_envs=num_envs, precision=precision)
    sim = env._sim
    E = num_envs
    PERIOD = 5
    NUM_WP = control_freq_hz * PERIOD
    TARGET_POS = np.zeros((NUM_WP, 2))
    for i in range(NUM_WP):
        TARGET_POS[i, :] = [0.5 * np.cos(2 * np.pi * (i / NUM_WP)), 0]
    wp = np.array([0, int(NUM_WP / 2)])
    ctrl = DSLPIDControl(drone_model=drone, num=E * 2, precision=precision)
    action = torch.zeros((E, 2, 4), dtype=sim.real, device=sim.device)
    zmin = np.inf
    for i in range(int(duration_sec * control_freq_hz)):
        obs, *_ = env.step(action)
        tp = np.array([[TARGET_POS[wp[j], 0], TARGET_POS[wp[j], 1], INIT_XYZS[j, 2]] for j in range(2)])
        tp_t = torch.as_tensor(np.tile(tp, (E, 1)), dtype=sim.real, device=sim.device)
        rpm, _, _ = ctrl.computeControlFromState(control_timestep=env.CTRL_TIMESTEP, state=obs.reshape(-1, 20), target_pos=tp_t)
        action = rpm.reshape(E, 2, 4)
        wp = np.where(wp < NUM_WP - 1, wp + 1, 0)
        zmin = min(zmin, float(obs[:, 1, 2].min()))
    print(f"[downwash] {E} envs: lowest altitude of the lower drone {zmin:.3f} m (starts at 0.5 m)")
    env.close()
    return zmin


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--num_envs", default=256, type=int)
    ap.add_argument("--precision", default="f64", choices=["f32", "f64"])
    a = ap.parse_args()
    run(num_envs=a.num_envs, precision=a.precision)
