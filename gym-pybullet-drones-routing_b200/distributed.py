"""Multi-GPU plumbing: independent envs are sharded by index across ranks (one process per GPU); nothing is
exchanged on the step path.  The only collective is the all-reduce of the episode statistics
(``gpd_episode_stats``), once per logging interval — what SB3's ``Monitor`` aggregates in the single-process
reference (``examples/learn.py:53-57,142-146``).  NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

STAT_NAMES = ("episodes", "sum_return", "sum_length", "sum_return_sq", "min_return", "max_return", "env_steps",
              "terminated_episodes")
_SUM_IDX = [0, 1, 2, 3, 6, 7]


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous env-index range of ``rank``: ``[start, start + count)``; an env never straddles ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def init_process_group_from_env(backend: str | None = None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, local, world


class NcclStatsComm:
    """An ``ncclComm_t`` owned by libgpd_b200 for the in-library statistics reduce (``gpd_episode_stats(..., comm, ...)``:
    one all-gather of 8 doubles + a rank-ordered combine on the device, instead of three ``torch.distributed`` all-reduces).
    ``torch.distributed`` does not expose its communicator, so rank 0 draws an NCCL unique id and the 128 bytes travel
    through whatever process group is up (NCCL or gloo); the communicator itself lives in the library.

        comm = NcclStatsComm(device=local_rank)            # after init_process_group
        job_stats = sim.episode_stats(nccl_comm=comm.handle)
    """

    def __init__(self, device: int, group=None):
        import ctypes as C

        from . import _lib
        self.lib = _lib.load()
        self.handle = None
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NcclStatsComm needs an initialised torch.distributed process group to exchange the unique id")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = C.create_string_buffer(128)
        if rank == 0:
            _lib.check(self.lib.gpd_nccl_unique_id(buf))
        dev = torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0, group=group)
        uid = bytes(t.cpu().tolist())
        h = C.c_void_p()
        _lib.check(self.lib.gpd_nccl_comm_init(uid, rank, world, int(device), C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle is not None:
            self.lib.gpd_nccl_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def all_reduce_episode_stats(stats, device=None, group=None) -> np.ndarray:
    """Job-wide statistics from per-rank ``gpd_episode_stats`` vectors: sums are added, min/max reduced.
    A rank with no finished episode contributes +inf/-inf to min/max."""
    s = np.asarray(stats, dtype=np.float64).copy()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return s
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    has = s[0] > 0
    sums = torch.tensor(s[_SUM_IDX], dtype=torch.float64, device=device)
    mn = torch.tensor([s[4] if has else np.inf], dtype=torch.float64, device=device)
    mx = torch.tensor([s[5] if has else -np.inf], dtype=torch.float64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out = np.zeros(8)
    out[_SUM_IDX] = sums.cpu().numpy()
    any_ep = out[0] > 0
    out[4] = float(mn.item()) if any_ep else 0.0
    out[5] = float(mx.item()) if any_ep else 0.0
    return out


def summarize(stats) -> dict:
    s = np.asarray(stats, dtype=np.float64)
    n = max(s[0], 1.0)
    mean = s[1] / n
    var = max(s[3] / n - mean * mean, 0.0)
    return {"episodes": int(s[0]), "ep_rew_mean": mean, "ep_rew_std": var ** 0.5, "ep_len_mean": s[2] / n,
            "ep_rew_min": s[4], "ep_rew_max": s[5], "env_steps": int(s[6]), "terminated_frac": s[7] / n}
