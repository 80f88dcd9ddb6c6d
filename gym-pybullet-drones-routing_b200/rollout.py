"""On-device rollouts: ``n_steps`` of (policy -> env.step) captured ONCE into a CUDA graph and replayed
(caller side: the PPO ``collect_rollouts`` loop behind ``examples/learn.py:84-94``).

At 65k envs a step kernel lasts ~11 us, about what the host needs to launch it: with the policy on the device the whole
rollout becomes one graph launch — no per-step host work, no host<->device copies, no Python between steps.
"""
from __future__ import annotations

from typing import Callable

import torch


def _rollout_chain(sim, policy, n_steps, obs, actions, rewards, terminated, truncated):
    """n_steps of (policy -> gpd_step) where the trajectory buffer IS the observation chain: step t reads the action ring
    from ``obs[t]`` and writes ``obs[t + 1]``, rewards and flags land in their trajectory rows — no per-step copies.  One
    copy in (the env's current observation) and one out (``adopt_obs``) per rollout keep the env usable eagerly."""
    obs[0].copy_(sim.obs)
    for t in range(n_steps):
        with torch.no_grad():
            a = policy(obs[t])
        actions[t].copy_(a.reshape(actions[t].shape))
        sim.step_into(actions[t], obs[t], obs[t + 1], rewards[t], terminated[t], truncated[t])
    sim.adopt_obs(obs[n_steps])


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
        _rollout_chain(self.sim, self.policy, self.n_steps, self.obs, self.actions, self.rewards, self.terminated, self.truncated)

    def run(self):
        """One rollout = one graph launch.  Returns views of the trajectory buffers (overwritten by the next run)."""
        self.env._state_cache = None
        self.graph.replay()
        return self.obs, self.actions, self.rewards, self.terminated.view(torch.bool), self.truncated.view(torch.bool)


class GraphedPoolRollout:
    """``GraphedRollout`` over several independent env pools: one CUDA graph whose branches (one per pool, each on its own
    captured stream) run ``n_steps`` of (policy -> env.step).  The branches overlap on the device — one pool's kernel
    boundary and store tail under another pool's loads (``pool.py``; 8 pools of 65,536 envs: 8.2 instead of 10.7 us per
    step) — and one ``run()`` is still a single graph launch.

    ``envs``    batched aviaries built with ``auto_reset=True`` (same shape); ``policy`` as in ``GraphedRollout``
    Trajectory buffers carry a leading pool axis: ``obs (P, n_steps + 1, E, N, W)``, ``actions (P, n_steps, E, N, A)`` ...
    """

    def __init__(self, envs, policy: Callable[[torch.Tensor], torch.Tensor], n_steps: int, warmup: int = 2):
        if n_steps < 2 or n_steps % 2:
            raise ValueError("n_steps must be even and >= 2 (observation ping-pong period)")
        self.envs, self.policy, self.n_steps = list(envs), policy, n_steps
        sims = [e._sim for e in self.envs]
        if not sims or any(not s.auto_reset for s in sims):
            raise ValueError("GraphedPoolRollout needs envs built with auto_reset=True")
        s0 = sims[0]
        if any((s.E, s.N, s.W, s.A) != (s0.E, s0.N, s0.W, s0.A) for s in sims):
            raise ValueError("all pools must have the same shape")
        dev, P = s0.device, len(sims)
        self.sims = sims
        self.obs = torch.zeros((P, n_steps + 1, s0.E, s0.N, s0.W), dtype=s0.obs_dtype, device=dev)
        self.actions = torch.zeros((P, n_steps, s0.E, s0.N, s0.A), dtype=s0.act_dtype, device=dev)
        self.rewards = torch.zeros((P, n_steps, s0.E), dtype=s0.real, device=dev)
        self.terminated = torch.zeros((P, n_steps, s0.E), dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros((P, n_steps, s0.E), dtype=torch.uint8, device=dev)
        self.branches = [torch.cuda.Stream(device=dev) for _ in range(P - 1)]
        for s in sims:
            if not s._have_prev:
                s.reset()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        phases = [s._cur for s in sims]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self._body()
        torch.cuda.synchronize(dev)
        assert phases == [s._cur for s in sims]
        self.graph = g

    def _branch(self, j):
        _rollout_chain(self.sims[j], self.policy, self.n_steps, self.obs[j], self.actions[j], self.rewards[j],
                       self.terminated[j], self.truncated[j])

    def _body(self):
        main = torch.cuda.current_stream(self.sims[0].device)
        fork = torch.cuda.Event()
        fork.record(main)
        for st in self.branches:
            st.wait_event(fork)
        for j in range(len(self.sims)):
            with torch.cuda.stream(main if j == 0 else self.branches[j - 1]):
                self._branch(j)
        for st in self.branches:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    def run(self):
        """One rollout of every pool = one graph launch.  Returns views of the trajectory buffers."""
        for e in self.envs:
            e._state_cache = None
        self.graph.replay()
        return self.obs, self.actions, self.rewards, self.terminated.view(torch.bool), self.truncated.view(torch.bool)
