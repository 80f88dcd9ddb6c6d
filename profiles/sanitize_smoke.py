#!/usr/bin/env python
"""Small driver for `compute-sanitizer --tool memcheck`: every step-kernel variant, reset, state export, PID, forces,
rollout — tiny shapes with ragged tails, 3 steps each."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import CtrlAviary, HoverAviary, MultiHoverAviary, VelocityAviary  # noqa: E402
from gpd_b200.utils.enums import ActionType, DroneModel, Physics  # noqa: E402

cases = []
for prec in ("f32", "f64"):
    cases += [
        ("hover rpm 30Hz", lambda p=prec: HoverAviary(num_envs=1000, precision=p, auto_reset=True), 4),
        ("hover rpm 48Hz", lambda p=prec: HoverAviary(num_envs=333, ctrl_freq=48, precision=p, auto_reset=True), 4),
        ("hover rpm 240Hz (B=120, no TMA)", lambda p=prec: HoverAviary(num_envs=70, ctrl_freq=240, precision=p), 4),
        ("hover pid 48Hz", lambda p=prec: HoverAviary(num_envs=333, ctrl_freq=48, act=ActionType.PID, drone_model=DroneModel.CF2P, precision=p, auto_reset=True), 3),
        ("hover 1d rpm 30Hz", lambda p=prec: HoverAviary(num_envs=130, act=ActionType.ONE_D_RPM, precision=p), 1),
        ("hover 1d pid 48Hz", lambda p=prec: HoverAviary(num_envs=130, ctrl_freq=48, act=ActionType.ONE_D_PID, precision=p), 1),
        ("hover vel 30Hz", lambda p=prec: HoverAviary(num_envs=130, act=ActionType.VEL, precision=p), 4),
        ("multihover 3 gnd+drag+dw", lambda p=prec: MultiHoverAviary(num_envs=77, num_drones=3, physics=Physics.DYN_GND_DRAG_DW, precision=p, auto_reset=True), 4),
        ("ctrl 64 dw", lambda p=prec: CtrlAviary(num_envs=5, num_drones=64, physics=Physics.DYN_DW, ctrl_freq=48, precision=p), 4),
        ("ctrl 256", lambda p=prec: CtrlAviary(num_envs=2, num_drones=256, ctrl_freq=48, precision=p), 4),
        ("velocity 2", lambda p=prec: VelocityAviary(num_envs=50, num_drones=2, ctrl_freq=48, drone_model=DroneModel.CF2P, precision=p), 4),
    ]
big = HoverAviary(num_envs=300000, precision="f32", auto_reset=True)      # whole-sector split path (>= 262,144 drones)
cases.append(("hover rpm 300k envs (edge boxes)", lambda: big, 4))
for name, mk, A in cases:
    env = mk()
    sim = env._sim
    env.reset()
    for t in range(3):
        if sim.is_ctrl and sim.action_type == "ctrl_rpm":
            a = torch.full((sim.E, sim.N, 4), float(env.HOVER_RPM), dtype=sim.real, device="cuda")
        else:
            a = (torch.rand((sim.E, sim.N, sim.A), device="cuda") * 2 - 1).to(sim.act_dtype)
        env.step(a)
    env.reset_envs(torch.rand(sim.E, device="cuda") < 0.5)
    st = sim.get_state()
    sim.set_state(*st)
    if sim.auto_reset:
        sim.episode_stats(clear=True)
    if sim.is_ctrl and sim.N == 2 and sim.action_type == "ctrl_rpm":
        pass
    torch.cuda.synchronize()
    print("ok", name, sim.precision, flush=True)
    env.close()
print("sanitize smoke done")
