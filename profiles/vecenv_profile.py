#!/usr/bin/env python
"""cProfile of GpdVecEnv.step(numpy) at 65,536 envs (1 and 2 pools): where the SB3-protocol bookkeeping spends its time."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import HoverAviary  # noqa: E402
from gpd_b200.vec_env import GpdVecEnv  # noqa: E402

E = 65536
rng = np.random.default_rng(0)
acts = [torch.from_numpy(rng.uniform(-1, 1, (E, 1, 4)).astype(np.float32)).pin_memory().numpy() for _ in range(4)]
for pools in (1, 2):
    venv = GpdVecEnv(HoverAviary, E, num_pools=pools, precision="f32")
    venv.reset()
    for k in range(10):
        venv.step(acts[k % 4])
    t0 = time.perf_counter()
    for k in range(200):
        o, r, d, infos = venv.step(acts[k % 4])
    dt = (time.perf_counter() - t0) / 200
    print(f"pools={pools}: {1e6 * dt:.1f} us per step")
    pr = cProfile.Profile()
    pr.enable()
    for k in range(200):
        venv.step(acts[k % 4])
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(8)
    venv.close()
