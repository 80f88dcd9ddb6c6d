#!/usr/bin/env python
"""Where the time of one numpy-facing step goes (HoverAviary.step(numpy) -> gpd_step_mirror_begin / _end), 65,536 envs."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import HoverAviary  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = HoverAviary(num_envs=E, auto_reset=True, precision="f32")
env.reset(as_numpy=True)
sim = env._sim
pin = [torch.empty((E, 1, 4), dtype=torch.float32).pin_memory() for _ in range(4)]
rng = np.random.default_rng(0)
for p in pin:
    p.copy_(torch.from_numpy(rng.uniform(-1, 1, (E, 1, 4)).astype(np.float32)))
acts = [p.numpy() for p in pin]
if os.environ.get("PAGEABLE_ACTIONS"):      # the library then stages the actions with a DMA copy and keeps zero-copy outputs
    acts = [np.array(a) for a in acts]
out = sim.alloc_host_outputs(pinned=True)
for k in range(20):
    sim.step_host(acts[k % 4], out)
n = 300
tb = te = tw = 0.0
for k in range(n):
    t0 = time.perf_counter()
    sim.step_host_begin(acts[k % 4], out)
    t1 = time.perf_counter()
    sim.step_host_end()
    t2 = time.perf_counter()
    tb += t1 - t0
    te += t2 - t1
# the host-side share of _end alone: let the device finish first, then complete the step
for k in range(50):
    sim.step_host_begin(acts[k % 4], out)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sim.step_host_end()
    tw += time.perf_counter() - t1
t0 = time.perf_counter()
for k in range(n):
    env.step(acts[k % 4])
tf = time.perf_counter() - t0
# the kernel alone with host I/O (device time): events around begin
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dev = 0.0
for k in range(50):
    e0.record()
    sim.step_host_begin(acts[k % 4], out)
    e1.record()
    sim.step_host_end()
    torch.cuda.synchronize()
    dev += e0.elapsed_time(e1)
print(json.dumps({"E": E, "zero_copy": os.environ.get("GPD_MIRROR_ZEROCOPY", "1"), "pageable_actions": bool(os.environ.get("PAGEABLE_ACTIONS")), "begin_us": 1e6 * tb / n, "end_us": 1e6 * te / n,
                  "end_host_only_us": 1e6 * tw / 50, "facade_step_us": 1e6 * tf / n, "device_us": 1e3 * dev / 50}))
