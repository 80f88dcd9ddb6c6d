"""On-device rollouts: ``n_steps`` of (policy -> env.step) captured ONCE into a CUDA graph and replayed
(caller side: the PPO ``collect_rollouts`` loop behind ``examples/learn.py:84-94``).

At 65k envs a step kernel lasts ~11 us, about what the host needs to launch it: with the policy on the device the whole
rollout becomes one graph launch — no per-step host work, no host<->device copies, no Python between steps.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedRollout:
    """Collects trajectories ``obs[t], action[t], reward[t], terminated[t], truncated[t]`` for ``t < n_steps``.

    ``env``     a batched aviary built with ``auto_reset=True`` (episodes restart inside the kernel)
    ``policy``  ``obs (E, N, W) float32 -> action (E, N, A)``; pure torch ops with static shapes (graph-capturable)
    ``n_steps`` must be even: the observation ping-pong of the env has period 2, so every replay sees the same buffers
    """

    def __init__(self, env, policy: Callable[[torch.Tensor], torch.Tensor], n_steps: int, warmup: int = 2):
        if n_steps < 2 or n_steps % 2:
            raise ValueError("n_steps must be even and >= 2 (observation ping-pong period)")
        sim = env._sim
        if not sim.auto_reset:
            raise ValueError("GraphedRollout needs an env built with auto_reset=True")
        self.env, self.sim, self.policy, self.n_steps = env, sim, policy, n_steps
        dev = sim.device
        E, N, W, A = sim.E, sim.N, sim.W, sim.A
        self.obs = torch.zeros((n_steps + 1, E, N, W), dtype=sim.obs_dtype, device=dev)
        self.actions = torch.zeros((n_steps, E, N, A), dtype=sim.act_dtype, device=dev)
        self.rewards = torch.zeros((n_steps, E), dtype=sim.real, device=dev)
        self.terminated = torch.zeros((n_steps, E), dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros((n_steps, E), dtype=torch.uint8, device=dev)
        self.graph = None
        if not sim._have_prev:
            sim.reset()
        # warm-up on a side stream (allocator, cuBLAS handles) — an even number of eager rollouts keeps the ping-pong phase
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._phase = sim._cur
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self._body()
        torch.cuda.synchronize(dev)
        assert sim._cur == self._phase
        self.graph = g

    def _body(self):
        sim = self.sim
        self.obs[0].copy_(sim.obs)
        for t in range(self.n_steps):
            with torch.no_grad():
                a = self.policy(self.obs[t])
            self.actions[t].copy_(a.reshape(self.actions[t].shape))
            o, r, te, tr = sim.step(self.actions[t])
            self.obs[t + 1].copy_(o)
            self.rewards[t].copy_(r)
            self.terminated[t].copy_(te)
            self.truncated[t].copy_(tr)

    def run(self):
        """One rollout = one graph launch.  Returns views of the trajectory buffers (overwritten by the next run)."""
        self.env._state_cache = None
        self.graph.replay()
        return self.obs, self.actions, self.rewards, self.terminated.view(torch.bool), self.truncated.view(torch.bool)
