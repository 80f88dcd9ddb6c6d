"""Host-side multi-rank logic on CPU: env sharding and the episode-statistics all-reduce (gloo, world_size 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gpd_b200  # noqa: F401
from gpd_b200.distributed import all_reduce_episode_stats, shard_range, summarize


@pytest.mark.parametrize("total,world", [(65536, 8), (4096, 3), (7, 8), (16777216, 8), (1, 1)])
def test_shard_range_partitions_envs(total, world):
    spans = [shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0
    assert sum(c for _, c in spans) == total
    for (s0, c0), (s1, _) in zip(spans, spans[1:]):
        assert s0 + c0 == s1                         # contiguous, no env straddles ranks
    counts = [c for _, c in spans]
    assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if rank == 0:
        local = np.array([3, 30.0, 45, 330.0, 5.0, 15.0, 100, 1], dtype=np.float64)
    else:
        local = np.array([0, 0.0, 0, 0.0, 0.0, 0.0, 50, 0], dtype=np.float64)      # no finished episode on this rank
    out = all_reduce_episode_stats(local)
    # env sharding: each rank handles its own slice; the union must equal the single-process result
    total = 37
    start, count = shard_range(total, rank, world)
    acts = np.random.default_rng(123).uniform(-1, 1, size=(total, 4))
    t = torch.tensor([acts[start:start + count].sum()], dtype=torch.float64)
    dist.all_reduce(t)
    q.put((rank, out.tolist(), float(t.item()), float(acts.sum())))
    dist.destroy_process_group()


def test_episode_stats_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, sharded_sum, full_sum in res:
        assert out[:4] == [3, 30.0, 45, 330.0]
        assert out[4] == 5.0 and out[5] == 15.0       # min/max ignore the rank without episodes
        assert out[6] == 150 and out[7] == 1
        assert abs(sharded_sum - full_sum) < 1e-12
    s = summarize(res[0][1])
    assert s["episodes"] == 3 and abs(s["ep_rew_mean"] - 10.0) < 1e-12 and abs(s["ep_len_mean"] - 15.0) < 1e-12


def test_all_reduce_is_identity_without_process_group():
    x = np.arange(8, dtype=np.float64)
    assert np.array_equal(all_reduce_episode_stats(x), x)
