import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import gpd_b200
from gpd_b200.envs import CtrlAviary
from gpd_b200.utils.enums import Physics
E, N = 4096, 64
rng = np.random.default_rng(1)
xyz = np.concatenate([rng.uniform(-2, 2, size=(E, N, 2)), rng.uniform(0.2, 3, size=(E, N, 1))], axis=-1)
env = CtrlAviary(num_envs=E, num_drones=N, physics=Physics.DYN_DW, pyb_freq=240, ctrl_freq=48, initial_xyzs=xyz, precision="f32")
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = [(env.HOVER_RPM * (1 + 0.02 * (torch.rand((E, N, 4), generator=g, device="cuda") * 2 - 1))).float() for _ in range(2)]
env.reset()
for k in range(6):
    env._sim.step(acts[k % 2])
torch.cuda.synchronize()
