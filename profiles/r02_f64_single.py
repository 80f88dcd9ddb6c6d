#!/usr/bin/env python
"""FP64 single-drone shapes that still run gpd::step_kernel (not the bulk kernel): CtrlAviary x1 and HoverAviary ONE_D_RPM at
30 Hz (27-float rows), plus the FP64 bulk shapes and C3, us per step in 64-step graphs over rotating env sets."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import CtrlAviary, HoverAviary, MultiHoverAviary  # noqa: E402
from gpd_b200.utils.enums import ActionType, Physics  # noqa: E402

timer = bench.Timer(torch, dist, 1, torch.device("cuda", 0))
g = torch.Generator(device="cuda"); g.manual_seed(0)
rand = lambda shape: (torch.rand(shape, generator=g, device="cuda") * 2 - 1)
E = 65536
out = {"label": os.environ.get("LABEL", "")}
cases = {
    "ctrl1_f64": (lambda: CtrlAviary(num_envs=E, num_drones=1, physics=Physics.DYN, ctrl_freq=48, precision="f64"),
                  lambda env, k: (env.HOVER_RPM * (1 + 0.05 * rand((E, 1, 4)))).double()),
    "ctrl1_gnd_drag_f64": (lambda: CtrlAviary(num_envs=E, num_drones=1, physics=Physics.DYN_GND_DRAG, ctrl_freq=48, precision="f64"),
                           lambda env, k: (env.HOVER_RPM * (1 + 0.05 * rand((E, 1, 4)))).double()),
    "hover_one_d_30hz_f64": (lambda: HoverAviary(num_envs=E, act=ActionType.ONE_D_RPM, ctrl_freq=30, precision="f64", auto_reset=True),
                             lambda env, k: rand((E, 1, 1))),
    "hover_pid_30hz_f64": (lambda: HoverAviary(num_envs=E, act=ActionType.PID, ctrl_freq=30, precision="f64", auto_reset=True),
                           lambda env, k: rand((E, 1, 3))),
    "hover_rpm_f64_bulk": (lambda: HoverAviary(num_envs=E, ctrl_freq=30, precision="f64", auto_reset=True),
                           lambda env, k: rand((E, 1, 4))),
    "hover_gnd_drag_f64_bulk": (lambda: HoverAviary(num_envs=E, physics=Physics.DYN_GND_DRAG, ctrl_freq=30, precision="f64", auto_reset=True),
                                lambda env, k: rand((E, 1, 4))),
    "c3_f64": (lambda: MultiHoverAviary(num_envs=32768, num_drones=2, physics=Physics.DYN_GND_DRAG, ctrl_freq=30, precision="f64", auto_reset=True),
               lambda env, k: rand((32768, 2, 4))),
}
for name, (mk, ma) in cases.items():
    r = bench.measure_config(torch, timer, mk, ma, nsets=6, steps=48)
    out[name] = round(r["us_per_step"], 2)
print(json.dumps(out))
