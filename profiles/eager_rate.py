"""Host cost of eager stepping (no CUDA graph): microseconds per BatchedSim.step / HoverAviary.step with device tensors."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpd_b200
from gpd_b200.envs import HoverAviary
for E in (4096, 65536):
    env = HoverAviary(num_envs=E, precision="f32", auto_reset=True)
    env.reset()
    a = torch.rand((E, 1, 4), device="cuda") * 2 - 1
    for _ in range(200): env._sim.step(a)
    torch.cuda.synchronize()
    for name, fn in (("sim.step", lambda: env._sim.step(a)), ("env.step", lambda: env.step(a))):
        t0 = time.perf_counter()
        for _ in range(5000): fn()
        torch.cuda.synchronize()
        print(E, name, round((time.perf_counter() - t0) / 5000 * 1e6, 2), "us per eager step")
