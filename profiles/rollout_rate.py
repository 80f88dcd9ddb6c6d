#!/usr/bin/env python
"""On-device rollouts (policy + env.step in one CUDA graph): microseconds per 65,536-env step, one pool vs several pools."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import HoverAviary  # noqa: E402
from gpd_b200.rollout import GraphedPoolRollout, GraphedRollout  # noqa: E402

E, T = 65536, 32
torch.manual_seed(0)
W1 = (0.05 * torch.randn(72, 64)).cuda()
W2 = (0.5 * torch.randn(64, 4)).cuda()


def policy(obs):
    return torch.tanh(torch.tanh(obs.reshape(obs.shape[0], -1) @ W1) @ W2).reshape(obs.shape[0], 1, 4)


def timed(ro, nsteps, reps=20):
    for _ in range(3):
        ro.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ro.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * nsteps) * 1e3


for P in (1, 2, 4, 8):
    envs = [HoverAviary(num_envs=E, auto_reset=True, precision="f32") for _ in range(P)]
    ro = GraphedRollout(envs[0], policy, T) if P == 1 else GraphedPoolRollout(envs, policy, T)
    us = timed(ro, T * P)
    print(json.dumps(dict(pools=P, envs_per_pool=E, us_per_pool_step=round(us, 2), env_steps_per_s=round(E / (us * 1e-6)),
                          note="MLP 72-64-4 policy + trajectory copies + env.step per step, one graph launch per rollout")), flush=True)
    del ro
    for e in envs:
        e.close()
    torch.cuda.empty_cache()
