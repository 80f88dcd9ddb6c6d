"""SB3-protocol ``VecEnv`` over one batched aviary (caller side: ``examples/learn.py:53-94``).

stable-baselines3 is not installable in the build image, so the adapter is duck-typed: ``num_envs``,
``observation_space``/``action_space`` (of ONE env), ``reset() -> obs``, ``step_async/step_wait``, ``step``,
``close``, ``get_attr/set_attr/env_method/env_is_wrapped/seed``, auto-reset with
``infos[i]["terminal_observation"]``, ``infos[i]["TimeLimit.truncated"]`` and Monitor-style
``infos[i]["episode"] = {"r","l","t"}``.  At 65k envs a Python list of dicts per step is the bottleneck, not the
kernel: ``infos`` is a lazy sequence that builds a dict only for the indices a consumer touches, and
``step_tensor`` is the tensor-native entry point for device-side policies.
"""
from __future__ import annotations

import time
from collections.abc import Sequence

import numpy as np
import torch


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
    def __init__(self, env_cls, num_envs: int, **env_kwargs):
        env_kwargs = dict(env_kwargs)
        env_kwargs.update(num_envs=num_envs, auto_reset=True)
        self.env = env_cls(**env_kwargs)
        self.num_envs = int(num_envs)
        self.observation_space = self.env.observation_space
        self.action_space = self.env.action_space
        self.render_mode = None
        self._actions = None
        self._t0 = time.time()
        self._ep_r = np.zeros(self.num_envs)
        self._ep_l = np.zeros(self.num_envs, dtype=np.int64)
        self.reset_infos = [{} for _ in range(min(self.num_envs, 1))]

    # ---- SB3 VecEnv protocol ----------------------------------------------------------
    def reset(self):
        obs, _ = self.env.reset(as_numpy=True)
        self._ep_r[:] = 0
        self._ep_l[:] = 0
        return obs

    def step_async(self, actions):
        self._actions = np.asarray(actions)

    def step_wait(self):
        sim = self.env._sim
        if self.env._host_out is None:
            self.env._host_out = sim.alloc_host_outputs(pinned=torch.cuda.is_available(), terminal_kin=True)
        obs, rew, term, trunc, tkin = sim.step_host(self._actions, self.env._host_out)
        self.env._state_cache = None
        term_b, trunc_b = term.view(np.bool_), trunc.view(np.bool_)
        dones = term_b | trunc_b
        self._ep_r += rew
        self._ep_l += 1
        infos = LazyInfos(self.num_envs, dones.copy(), term_b.copy(), trunc_b.copy(), obs, tkin,
                          self._ep_r.copy(), self._ep_l.copy(), self._t0)
        self._ep_r[dones] = 0
        self._ep_l[dones] = 0
        return obs, rew, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_tensor(self, actions: torch.Tensor):
        """Device-side step: CUDA tensors in and out, no host copies, no info dicts.
        Returns (obs, reward, terminated, truncated); ``self.env._sim.terminal_kin`` holds the terminal rows."""
        self.env._state_cache = None
        obs, rew, term, trunc = self.env._sim.step(actions)
        return obs, rew, term.view(torch.bool), trunc.view(torch.bool)

    def close(self):
        self.env.close()

    def seed(self, seed=None):
        return [None] * self.num_envs      # the reference ignores seeds too (BaseAviary.py:243)

    def get_attr(self, attr_name, indices=None):
        v = getattr(self.env, attr_name)
        return [v for _ in self._idx(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        r = getattr(self.env, method_name)(*args, **kwargs)
        return [r for _ in self._idx(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._idx(indices)]

    def episode_stats(self, clear=False):
        return self.env._sim.episode_stats(clear)

    def _idx(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices
