"""Batched counterpart of the reference's ``examples/pid.py``: ``num_envs`` copies of the circle-tracking demo
(CtrlAviary + one DSLPIDControl per drone, 240 Hz sim / 48 Hz ctrl) on ``Physics.DYN``.

    python -m gpd_b200.examples.pid --num_envs 4096 --num_drones 3 --drone cf2p [--fused] [--log]

``--fused`` runs the whole loop inside one kernel launch per chunk (``gpd_rollout_pid``) instead of calling
``env.step`` + ``ctrl.computeControlFromState`` from Python.  Note (reference quirk, SURVEY finding 5): with the
default cf2x on Physics.DYN the reference's own demo flips the drone within ~14 control steps; cf2p tracks the circle.
"""
import argparse
import time

import numpy as np
import torch

from ..control.DSLPIDControl import DSLPIDControl
from ..envs.CtrlAviary import CtrlAviary
from ..params import default_pid_params
from ..utils.enums import DroneModel, Physics
from ..utils.Logger import Logger


def run(drone=DroneModel("cf2p"), num_drones=3, num_envs=1024, physics=Physics("dyn"), simulation_freq_hz=240,
        control_freq_hz=48, duration_sec=12, fused=False, log=False, output_folder="results", precision="f64", device=0):
    H, H_STEP, R = .1, .05, .3
    INIT_XYZS = np.array([[R * np.cos((i / 6) * 2 * np.pi + np.pi / 2), R * np.sin((i / 6) * 2 * np.pi + np.pi / 2) - R,
                           H + i * H_STEP] for i in range(num_drones)])
    INIT_RPYS = np.array([[0, 0, i * (np.pi / 2) / num_drones] for i in range(num_drones)])
    PERIOD = 10
    NUM_WP = control_freq_hz * PERIOD
    TARGET_POS = np.zeros((NUM_WP, 3))
    for i in range(NUM_WP):
        TARGET_POS[i, :] = (R * np.cos((i / NUM_WP) * (2 * np.pi) + np.pi / 2) + INIT_XYZS[0, 0],
                            R * np.sin((i / NUM_WP) * (2 * np.pi) + np.pi / 2) - R + INIT_XYZS[0, 1], 0)
    wp0 = np.array([int((i * NUM_WP / 6) % NUM_WP) for i in range(num_drones)])

    env = CtrlAviary(drone_model=drone, num_drones=num_drones, initial_xyzs=INIT_XYZS, initial_rpys=INIT_RPYS,
                     physics=physics, neighbourhood_radius=10, pyb_freq=simulation_freq_hz, ctrl_freq=control_freq_hz,
                     num_envs=num_envs, precision=precision, device=device)
    sim = env._sim
    real = sim.real
    E, N = num_envs, num_drones
    steps = int(duration_sec * env.CTRL_FREQ)
    wps = torch.as_tensor(TARGET_POS, dtype=real, device=sim.device)
    wp = torch.as_tensor(np.tile(wp0[None], (E, 1)), dtype=torch.int32, device=sim.device).contiguous()
    action = torch.zeros((E, N, 4), dtype=real, device=sim.device)
    logger = Logger(logging_freq_hz=control_freq_hz, num_drones=N, output_folder=output_folder) if log else None
    torch.cuda.synchronize()
    t0 = time.time()
    if fused:
        # the controller of gpd_rollout_pid is the one given to the simulator: build it with the drone's own model
        env.close()
        from ..sim import BatchedSim
        sim = BatchedSim(env.PARAMS, E, N, env_kind="ctrl", action_type="ctrl_rpm", pyb_freq=simulation_freq_hz,
                         ctrl_freq=control_freq_hz, precision=precision, device=device, pid=default_pid_params(drone),
                         init_xyz=INIT_XYZS, init_rpy=INIT_RPYS)
        chunk = control_freq_hz
        for s0 in range(0, steps, chunk):
            sim.rollout_pid(min(chunk, steps - s0), wps, wp, action)
            if logger is not None:
                logger.log_batch((s0 + chunk) / control_freq_hz, sim.get_state()[0][0])
    else:
        ctrl = DSLPIDControl(drone_model=drone, num=E * N, device=device, precision=precision)
        init_z = torch.as_tensor(INIT_XYZS[:, 2], dtype=real, device=sim.device).repeat(E)
        trpy = torch.as_tensor(INIT_RPYS, dtype=real, device=sim.device).repeat(E, 1)
        wpl = wp.long()
        for i in range(steps):
            obs, reward, terminated, truncated, info = env.step(action)
            tp = torch.cat([wps[wpl.reshape(-1)][:, 0:2], init_z[:, None]], dim=1)
            rpm, _, _ = ctrl.computeControlFromState(control_timestep=env.CTRL_TIMESTEP, state=obs.reshape(-1, 20),
                                                     target_pos=tp, target_rpy=trpy)
            action = rpm.reshape(E, N, 4)
            wpl = torch.where(wpl < NUM_WP - 1, wpl + 1, torch.zeros_like(wpl))
            if logger is not None:
                ctl = np.hstack([TARGET_POS[wpl[0].cpu().numpy(), 0:2], INIT_XYZS[:, 2:3], INIT_RPYS, np.zeros((N, 6))])
                logger.log_batch(i / env.CTRL_FREQ, obs[0], ctl)
    torch.cuda.synchronize()
    wall = time.time() - t0
    st = sim.get_state()[0]
    err = torch.linalg.norm(st[..., 0:2] - wps[(wp.long() if fused else wpl)][..., 0:2], dim=-1)
    print(f"[pid] {E} envs x {N} drones, {steps} ctrl steps in {wall:.2f} s "
          f"({E * N * steps * sim.S / wall:.3g} drone-substeps/s incl. Python); "
          f"median xy tracking error {float(err.median()):.4f} m, z range [{float(st[..., 2].min()):.3f}, {float(st[..., 2].max()):.3f}]")
    if logger is not None:
        logger.save()
        logger.save_as_csv("pid")
    sim.close()
    return float(err.median())


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Batched circle-tracking PID demo on Physics.DYN")
    ap.add_argument("--drone", default="cf2p", type=DroneModel, choices=[DroneModel.CF2X, DroneModel.CF2P])
    ap.add_argument("--num_drones", default=3, type=int)
    ap.add_argument("--num_envs", default=1024, type=int)
    ap.add_argument("--duration_sec", default=12, type=int)
    ap.add_argument("--precision", default="f64", choices=["f32", "f64"])
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--log", action="store_true")
    a = ap.parse_args()
    run(drone=a.drone, num_drones=a.num_drones, num_envs=a.num_envs, duration_sec=a.duration_sec, fused=a.fused, log=a.log,
        precision=a.precision)
