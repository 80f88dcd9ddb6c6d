#!/usr/bin/env python
"""Per-CTA phase timeline of the step kernel (gpd_set_timeline_buffer). Prints percentiles of each phase relative to
the earliest CTA start of a launch, for a launch in the middle of a graph-free sequence."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa: E402,F401
from gpd_b200 import _lib  # noqa: E402
from gpd_b200.envs import HoverAviary  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tpb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nsets = int(sys.argv[3]) if len(sys.argv) > 3 else 8
envs = [HoverAviary(num_envs=E, auto_reset=True, threads_per_block=tpb) for _ in range(nsets)]
L = _lib.load()
grid = L.gpd_grid_size(envs[0]._sim.h)
bufs = [torch.zeros((grid, 8), dtype=torch.int64, device="cuda") for _ in range(nsets)]
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [(torch.rand((E, 1, 4), generator=g, device="cuda") * 2 - 1) for _ in range(2 * nsets)]
for e in envs:
    e.reset()
    e._sim.set_step_chaining(True)      # pre-generated actions: launches overlap across the kernel boundary
for rep in range(4):
    for k in range(2 * nsets):
        envs[k % nsets]._sim.step(acts[k])
torch.cuda.synchronize()
for e, b in zip(envs, bufs):
    _lib.check(L.gpd_set_timeline_buffer(e._sim.h, C.c_void_p(b.data_ptr())))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for k in range(2 * nsets):
            envs[k % nsets]._sim.step(acts[k])
torch.cuda.synchronize()
for _ in range(5):
    gr.replay()
torch.cuda.synchronize()
names = ["cta_start", "pdl_wait_done", "state_arrived", "substeps_done", "phys_stored", "tma_loaded", "tma_store_read", "barrier"]
if os.environ.get("GPD_BULK", "1") != "0":     # phases of gpd::step_kernel_bulk (single-drone RL shapes)
    names = ["cta_start", "tile_claimed", "tile_loaded", "substeps_done", "tile_patched", "-", "-", "tile_published"]
# the last replay wrote each buffer twice (period 2*nsets); look at set 3's last launch and its predecessor (set 2)
t3 = bufs[3 % nsets].cpu().numpy().astype(np.int64)
t2 = bufs[2 % nsets].cpu().numpy().astype(np.int64)
t0 = t3[:, 0].min()
print(f"grid={grid} E={E}; times in us relative to the first CTA start of this launch")
print(f"previous launch (other env set): last barrier at {(t2[:, 7].max() - t0) / 1e3:+.2f} us, its first CTA start at {(t2[:, 0].min() - t0) / 1e3:+.2f} us")
for j, n in enumerate(names):
    v = (t3[:, j] - t0) / 1e3
    v = v[t3[:, j] > 0]
    if len(v):
        print(f"{n:16s} min {v.min():7.2f}  p10 {np.percentile(v, 10):7.2f}  p50 {np.percentile(v, 50):7.2f}  p90 {np.percentile(v, 90):7.2f}  max {v.max():7.2f}")
