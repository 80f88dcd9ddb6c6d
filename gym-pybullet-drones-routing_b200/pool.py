"""Asynchronous env pools: several independent batched aviaries, each stepping on its own CUDA stream.

One ``env.step()`` of 65,536 envs is a single wave of CTAs: while its last blocks compute and store, and across the kernel
boundary, DRAM idles.  Independent env sets have no such dependency on each other, so a trainer that keeps several pools
(what ``SubprocVecEnv`` workers are in the single-process reference, ``examples/learn.py:53-57``) can overlap them:
measured 10.7 -> 8.2 us per 65,536-env step with 8 pools (``python bench.py --streams 8``, profiles/README.md).

Each pool's own steps stay ordered (they run on the pool's stream); ``wait(j)`` orders the CALLER's stream behind pool
``j``'s latest step before its outputs are read.  Nothing here synchronises the host.
"""
from __future__ import annotations

import torch


class AsyncEnvPools:
    """``pools = AsyncEnvPools([env0, env1, ...])``; ``pools.step_async(j, action)``; ``obs, r, te, tr = pools.wait(j)``."""

    def __init__(self, envs, streams=None):
        if not envs:
            raise ValueError("AsyncEnvPools needs at least one env")
        self.envs = list(envs)
        dev = self.envs[0]._sim.device
        self.device = dev
        self.streams = list(streams) if streams is not None else [torch.cuda.Stream(device=dev) for _ in self.envs]
        if len(self.streams) != len(self.envs):
            raise ValueError("one stream per env pool")
        self._done = [None] * len(self.envs)
        self._out = [None] * len(self.envs)

    def __len__(self):
        return len(self.envs)

    def reset(self):
        """Resets every pool on its stream; returns the list of initial observations (ordered behind the caller's stream)."""
        cur = torch.cuda.current_stream(self.device)
        obs = []
        for j, (env, st) in enumerate(zip(self.envs, self.streams)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                o, _ = env.reset()
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)
            self._done[j], self._out[j] = None, None
            obs.append(o)
        return obs

    def step_async(self, j: int, action: torch.Tensor):
        """Launches pool ``j``'s step on its stream, after everything the caller's stream has enqueued so far (the action)."""
        st = self.streams[j]
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            self._out[j] = self.envs[j]._sim.step(action)
            ev = torch.cuda.Event()
            ev.record(st)
        action.record_stream(st)
        self._done[j] = ev

    def wait(self, j: int):
        """Orders the caller's stream behind pool ``j``'s latest step and returns ``(obs, reward, terminated, truncated)``."""
        if self._done[j] is None:
            raise RuntimeError(f"pool {j}: wait() without a step_async() in flight")
        torch.cuda.current_stream(self.device).wait_event(self._done[j])
        out, self._done[j] = self._out[j], None
        return out

    def step_all(self, actions):
        """Steps every pool (async), then waits for all: the outputs of ``[env.step(a) for env, a in zip(envs, actions)]``."""
        for j, a in enumerate(actions):
            self.step_async(j, a)
        return [self.wait(j) for j in range(len(self.envs))]

    def close(self):
        for e in self.envs:
            e.close()
