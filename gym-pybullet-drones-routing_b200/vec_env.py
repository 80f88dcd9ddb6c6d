"""SB3-protocol ``VecEnv`` over one batched aviary or several env pools (caller side:
``examples/learn.py:53-94``).

stable-baselines3 is not installable in the build image, so the adapter is duck-typed: ``num_envs``,
``observation_space``/``action_space`` (of ONE env), ``reset() -> obs``, ``step_async/step_wait``, ``step``,
``close``, ``get_attr/set_attr/env_method/env_is_wrapped/seed``, auto-reset with
``infos[i]["terminal_observation"]``, ``infos[i]["TimeLimit.truncated"]`` and Monitor-style
``infos[i]["episode"] = {"r","l","t"}``.  At 65k envs a Python list of dicts per step is the bottleneck, not the
kernel: ``infos`` is a lazy sequence that builds a dict only for the indices a consumer touches, and
``step_tensor`` is the tensor-native entry point for device-side policies.

``num_pools > 1`` splits the envs into equal pools (what ``SubprocVecEnv`` workers are in the reference,
``examples/learn.py:53-57``), each a batched aviary stepping on its own CUDA stream.  One numpy batch goes in and one
comes out: the pools share ONE host mirror (``sim.HostMirror``, a pinned feature-major log) through different column
offsets, so the observation is a single strided view over all pools, and because PCIe is full duplex one pool's
host-to-device action copy overlaps another pool's device-to-host result copy.
"""
from __future__ import annotations

import contextlib
import time
from collections.abc import Sequence

import numpy as np
import torch

from .sim import HostMirror


class LazyInfos(Sequence):
    def __init__(self, n, done, terminated, truncated, obs, terminal_kin, ep_r, ep_l, t0):
        self._n, self._done, self._term, self._trunc = n, done, terminated, truncated
        self._obs, self._tkin, self._ep_r, self._ep_l, self._t0 = obs, terminal_kin, ep_r, ep_l, t0

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        info = {"answer": 42, "TimeLimit.truncated": bool(self._trunc[i] and not self._term[i])}
        if self._done[i]:
            term_obs = np.array(self._obs[i], copy=True)      # ring part survives the reset (BaseRLAviary.py:153-154)
            if self._tkin is not None:
                term_obs[:, :12] = self._tkin[i]
            info["terminal_observation"] = term_obs
            info["episode"] = {"r": float(self._ep_r[i]), "l": int(self._ep_l[i]), "t": round(time.time() - self._t0, 6)}
        return info


class GpdVecEnv:
    def __init__(self, env_cls, num_envs: int, num_pools: int = 1, slide_steps: int = 128, **env_kwargs):
        env_kwargs = dict(env_kwargs)
        self.num_envs = int(num_envs)
        self.num_pools = int(num_pools)
        if self.num_pools < 1 or self.num_envs % self.num_pools:
            raise ValueError("num_envs must be a multiple of num_pools")
        per = self.num_envs // self.num_pools
        env_kwargs.update(num_envs=per, auto_reset=True)
        self.envs = [env_cls(**env_kwargs) for _ in range(self.num_pools)]
        self.env = self.envs[0]
        self.observation_space = self.env.observation_space
        self.action_space = self.env.action_space
        self.render_mode = None
        self._actions = None
        self._t0 = time.time()
        # running episode return / length of every env, as SB3's VecMonitor keeps them (in the sim's Real; int32).  Two buffers:
        # a step's infos read the one the step accumulated into, the other receives the zeroed continuation — like the
        # observation view, the infos of a step are valid until the next step() (SB3 consumes them right after step_wait)
        rdt0 = np.float64 if env_kwargs.get("precision", "f32") == "f64" else np.float32
        self._ep_r = [np.zeros(self.num_envs, rdt0), np.zeros(self.num_envs, rdt0)]
        self._ep_l = [np.zeros(self.num_envs, np.int32), np.zeros(self.num_envs, np.int32)]
        self._epc = 0
        self.reset_infos = [{} for _ in range(min(self.num_envs, 1))]
        sim0 = self.env._sim
        self._per, self._N = per, sim0.N
        self._device = sim0.device
        self._streams = [None] + [torch.cuda.Stream(device=self._device) for _ in range(self.num_pools - 1)]
        self._mirror = None
        self._outs = None
        self._slide_steps = int(slide_steps)

    # ---- host buffers: ONE mirror and ONE set of result arrays for all pools --------------
    def _ensure_host(self):
        if self._outs is not None:
            return
        sim0 = self.env._sim
        E, N, P = self.num_envs, self._N, self.num_pools
        if not sim0.is_ctrl:
            self._mirror = HostMirror(sim0.W, sim0.A, E * N, slide_steps=self._slide_steps)
            for j, env in enumerate(self.envs):
                env._sim.attach_mirror(self._mirror, col0=j * self._per * N)
        rdt = np.dtype(sim0.np_real)
        pin = torch.cuda.is_available()

        def host(shape, dt):
            n = int(np.prod(shape)) * np.dtype(dt).itemsize
            b = torch.empty(n, dtype=torch.uint8).pin_memory().numpy() if pin else np.empty(n, np.uint8)
            return b.view(dt).reshape(shape)
        # per pool [reward | terminated | truncated] packed like the device staging: one device-to-host copy per pool
        blk = self._per * (rdt.itemsize + 2)
        raw = host((P, blk), np.uint8)
        self._rew_p = [raw[j, :self._per * rdt.itemsize].view(rdt) for j in range(P)]
        self._term_p = [raw[j, self._per * rdt.itemsize:self._per * (rdt.itemsize + 1)] for j in range(P)]
        self._trunc_p = [raw[j, self._per * (rdt.itemsize + 1):] for j in range(P)]
        self._tkin = host((E, N, 12), np.float32) if not sim0.is_ctrl else None
        self._rew = np.empty(E, rdt)
        self._term = np.empty(E, np.bool_)
        self._trunc = np.empty(E, np.bool_)
        self._obs_ctrl = host((E, N, sim0.W), rdt) if sim0.is_ctrl else None
        self._outs = [(None if self._obs_ctrl is None else self._obs_ctrl[j * self._per:(j + 1) * self._per],
                       self._rew_p[j], self._term_p[j], self._trunc_p[j],
                       None if self._tkin is None else self._tkin[j * self._per:(j + 1) * self._per]) for j in range(P)]

    def _on(self, j):
        st = self._streams[j]       # pool 0 stays on the caller's current stream: no context switch at all (torch's costs ~10 us)
        return torch.cuda.stream(st) if st is not None else contextlib.nullcontext()

    # ---- SB3 VecEnv protocol ----------------------------------------------------------
    def reset(self):
        self._ensure_host()
        obs = None
        for j, env in enumerate(self.envs):
            with self._on(j):
                o, _ = env.reset(as_numpy=True)
            if self._obs_ctrl is not None:
                self._obs_ctrl[j * self._per:(j + 1) * self._per] = o
        if self._mirror is not None:
            obs = self._mirror.view(self.env._sim._row.value, self.num_envs, self._N, 0)
        else:
            obs = self._obs_ctrl
        for b in self._ep_r + self._ep_l:
            b[:] = 0
        return obs

    def step_async(self, actions):
        self._actions = np.asarray(actions)

    def step_wait(self):
        self._ensure_host()
        P, per = self.num_pools, self._per
        acts = self._actions.reshape(self.num_envs, -1)
        if self._mirror is not None:
            for j, env in enumerate(self.envs):            # enqueue every pool, then complete them in order
                with self._on(j):
                    env._sim.step_host_begin(acts[j * per:(j + 1) * per], self._outs[j])
                env._state_cache = None
            for j, env in enumerate(self.envs):
                with self._on(j):
                    row = env._sim.step_host_end()
            obs = self._mirror.view(row, self.num_envs, self._N, 0)
        else:
            for j, env in enumerate(self.envs):
                with self._on(j):
                    env._sim.step_host(acts[j * per:(j + 1) * per], self._outs[j])
                env._state_cache = None
            obs = self._obs_ctrl
        if P == 1:              # the pinned result block itself (valid until the next step, like the observation view)
            rew, term_b, trunc_b = self._rew_p[0], self._term_p[0].view(np.bool_), self._trunc_p[0].view(np.bool_)
        else:
            for j in range(P):
                sl = slice(j * per, (j + 1) * per)
                self._rew[sl] = self._rew_p[j]
                self._term[sl] = self._term_p[j].view(np.bool_)
                self._trunc[sl] = self._trunc_p[j].view(np.bool_)
            rew, term_b, trunc_b = self._rew, self._term, self._trunc
        # episode bookkeeping in whole-array passes (a boolean-mask assignment alone costs 230 us at 65,536 envs)
        c, n = self._epc, self._epc ^ 1
        er, el = self._ep_r[c], self._ep_l[c]
        np.add(er, rew, out=er)
        np.add(el, 1, out=el)
        dones = np.logical_or(term_b, trunc_b)
        keep = np.logical_not(dones)
        np.multiply(er, keep, out=self._ep_r[n])
        np.multiply(el, keep, out=self._ep_l[n])
        self._epc = n
        infos = LazyInfos(self.num_envs, dones, term_b.copy(), trunc_b.copy(), obs, self._tkin, er, el, self._t0)
        return obs, rew, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_tensor(self, actions: torch.Tensor):
        """Device-side step: CUDA tensors in and out, no host copies, no info dicts.
        Returns (obs, reward, terminated, truncated); ``self.env._sim.terminal_kin`` holds the terminal rows.
        With several pools the per-pool outputs are concatenated (one device copy each)."""
        if self.num_pools == 1:
            self.env._state_cache = None
            obs, rew, term, trunc = self.env._sim.step(actions)
            return obs, rew, term.view(torch.bool), trunc.view(torch.bool)
        per = self._per
        cur = torch.cuda.current_stream(self._device)
        outs = []
        for j, env in enumerate(self.envs):
            env._state_cache = None
            st = self._streams[j]
            if st is not None:
                st.wait_stream(cur)
            with self._on(j):
                outs.append(env._sim.step(actions[j * per:(j + 1) * per]))
        for st in self._streams:
            if st is not None:
                cur.wait_stream(st)
        obs, rew, term, trunc = (torch.cat([o[k] for o in outs]) for k in range(4))
        return obs, rew, term.view(torch.bool), trunc.view(torch.bool)

    def close(self):
        for e in self.envs:
            e.close()

    def seed(self, seed=None):
        return [None] * self.num_envs      # the reference ignores seeds too (BaseAviary.py:243)

    def get_attr(self, attr_name, indices=None):
        v = getattr(self.env, attr_name)
        return [v for _ in self._idx(indices)]

    def set_attr(self, attr_name, value, indices=None):
        for e in self.envs:
            setattr(e, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        r = getattr(self.env, method_name)(*args, **kwargs)
        return [r for _ in self._idx(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._idx(indices)]

    def episode_stats(self, clear=False):
        s = np.zeros(8)
        any_ep = False
        for e in self.envs:
            v = e._sim.episode_stats(clear)
            s[[0, 1, 2, 3, 6, 7]] += v[[0, 1, 2, 3, 6, 7]]
            if v[0] > 0:
                s[4] = v[4] if not any_ep else min(s[4], v[4])
                s[5] = v[5] if not any_ep else max(s[5], v[5])
                any_ep = True
        return s

    def _idx(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices
