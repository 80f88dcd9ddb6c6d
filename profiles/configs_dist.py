#!/usr/bin/env python
"""Multi-GPU protocol of SURVEY 8e for the non-headline shapes (one process per GPU, env-sharded, no data-path collective):
C5 weak scaling (2,097,152 envs per GPU, DSLPID in the loop) and C4 strong scaling (4,096 envs x 64 drones split over
the ranks).  Times are CUDA events, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/configs_dist.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import gpd_b200  # noqa: E402,F401
from gpd_b200.distributed import shard_range  # noqa: E402
from gpd_b200.envs import CtrlAviary, HoverAviary  # noqa: E402
from gpd_b200.utils.enums import ActionType, DroneModel, Physics  # noqa: E402
from configs import measure, rand_act  # noqa: E402


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def agg(us):
        t = torch.tensor([us], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # C5, weak: 2,097,152 envs per GPU
    E = 2097152
    r = measure(lambda: HoverAviary(num_envs=E, drone_model=DroneModel.CF2P, ctrl_freq=48, act=ActionType.PID, precision="f32",
                                    auto_reset=True, device=local), lambda env, k: rand_act((E, 1, 3), k + 100 * rank), nsets=1, reps=6)
    us = agg(r["us_per_step"])
    if rank == 0:
        print(json.dumps(dict(config="c5_hover_pid_48hz_f32", scaling="weak", n_gpus=world, envs_per_gpu=E, us_per_step=round(us, 1),
                              drone_substeps_per_s=world * E * 5 / (us * 1e-6))), flush=True)

    # C4, strong: 4,096 envs x 64 drones in total
    Etot, N = 4096, 64
    lo, El = shard_range(Etot, rank, world)
    hi = lo + El
    rng = np.random.default_rng(1)
    xyz = np.concatenate([rng.uniform(-2, 2, size=(Etot, N, 2)), rng.uniform(0.2, 3, size=(Etot, N, 1))], axis=-1)[lo:hi]

    def mk():
        return CtrlAviary(num_envs=El, num_drones=N, physics=Physics.DYN_DW, pyb_freq=240, ctrl_freq=48, initial_xyzs=xyz,
                          precision="f32", device=local)
    r = measure(mk, lambda env, k: (env.HOVER_RPM * (1 + 0.02 * rand_act((El, N, 4), k))).float(), nsets=4, reps=10)
    us = agg(r["us_per_step"])
    if rank == 0:
        print(json.dumps(dict(config="c4_ctrl64_dw_f32", scaling="strong", n_gpus=world, envs_per_gpu=El, threads_per_gpu=El * N,
                              us_per_step=round(us, 1), drone_substeps_per_s=Etot * N * 5 / (us * 1e-6))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
