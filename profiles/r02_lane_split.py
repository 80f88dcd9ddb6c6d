#!/usr/bin/env python
"""C4 shape (CtrlAviary, 64 drones per env, DYN + O(N^2) downwash, FP32, 240/48) at the per-rank sizes of the strong-scaling
run (4,096 envs in total over 1/2/4/8 GPUs): us per step with 1, 2 and 4 lanes per drone (GPD_LANE_SPLIT) and the default."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import CtrlAviary  # noqa: E402
from gpd_b200.utils.enums import Physics  # noqa: E402

N = 64
rng = np.random.default_rng(1)
for E in (512, 1024, 2048, 4096):
    xyz = np.concatenate([rng.uniform(-2, 2, size=(E, N, 2)), rng.uniform(0.2, 3, size=(E, N, 1))], axis=-1)
    for split in ("1", "2", "4", "auto"):
        if split == "auto":
            os.environ.pop("GPD_LANE_SPLIT", None)
        else:
            os.environ["GPD_LANE_SPLIT"] = split
        nsets = 4
        envs = [CtrlAviary(num_envs=E, num_drones=N, physics=Physics.DYN_DW, pyb_freq=240, ctrl_freq=48, initial_xyzs=xyz,
                           precision="f32") for _ in range(nsets)]
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        acts = [(envs[0].HOVER_RPM * (1 + 0.02 * (torch.rand((E, N, 4), generator=g, device="cuda") * 2 - 1))).float() for _ in range(4)]
        for e in envs:
            e.reset()
            e._sim.set_step_chaining(True)
        def run(n):
            for k in range(n):
                envs[k % nsets]._sim.step(acts[k % 4])
        run(8)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                run(32)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 32)
        print(json.dumps({"E": E, "N": N, "lanes": split, "us_per_step": round(1e3 * best, 2),
                          "pair_evals_per_s": E * N * N * 5 / (best * 1e-3)}), flush=True)
        for e in envs:
            e.close()
        del envs, gr
        torch.cuda.empty_cache()
