"""Times the UNMODIFIED reference Python (staged under oracle/_ref by oracle/stage_reference.py) on the host cores.
TEST/BENCH INFRASTRUCTURE: imported only by bench.py's reference arm / cpu_baseline leg and by tests/.

The reference's own ``HoverAviary(physics=Physics.DYN)`` runs under the pybullet/gymnasium stand-ins of oracle/refshim
(pybullet is not installable offline; on ``Physics.DYN`` Bullet only stores and returns the base pose, BaseAviary.py:862-872).
Scaling follows SB3 ``SubprocVecEnv`` (examples/learn.py:53-57 builds the vec env): one worker process per
host core, each stepping its own env instances and resetting an env when its episode ends.  Workers free-run between a
common start and their own finish (no per-step pipe round trip), so the figure is an upper bound of what SubprocVecEnv
delivers on the same cores.
"""
from __future__ import annotations

import contextlib
import io
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "refshim")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "gym_pybullet_drones"))


def _load():
    warnings.filterwarnings("ignore")
    if SHIM not in sys.path:
        sys.path[:0] = [SHIM, REF_ROOT]
    with contextlib.redirect_stdout(io.StringIO()):
        from gym_pybullet_drones.envs.HoverAviary import HoverAviary
        from gym_pybullet_drones.utils.enums import ActionType, ObservationType, Physics
    return HoverAviary, Physics, ObservationType, ActionType


def _worker(wid, n_envs, warmup, steps, ctrl_freq, seed, start, q):
    try:
        HoverAviary, Physics, ObservationType, ActionType = _load()
        with contextlib.redirect_stdout(io.StringIO()):
            envs = [HoverAviary(physics=Physics.DYN, ctrl_freq=ctrl_freq, obs=ObservationType.KIN, act=ActionType.RPM)
                    for _ in range(n_envs)]
            for e in envs:
                e.reset()
        rng = np.random.default_rng(seed + wid)

        def vec_step():
            acts = rng.uniform(-1, 1, size=(n_envs, 1, 4)).astype(np.float32)
            done = 0
            for k, e in enumerate(envs):
                _, _, te, tr, _ = e.step(acts[k])
                if te or tr:
                    e.reset()
                    done += 1
            return done
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(warmup):
                vec_step()
            start.wait()
            t0 = time.perf_counter()
            resets = 0
            for _ in range(steps):
                resets += vec_step()
            el = time.perf_counter() - t0
        q.put((wid, el, resets, None))
    except Exception as ex:        # never hang the parent
        try:
            start.abort()
        except Exception:
            pass
        q.put((wid, 0.0, 0, repr(ex)))


def calibrate(ctrl_freq=30, n=4, steps=3):
    """Seconds per env.step() of one reference env on one core (first-touch costs excluded)."""
    HoverAviary, Physics, ObservationType, ActionType = _load()
    with contextlib.redirect_stdout(io.StringIO()):
        envs = [HoverAviary(physics=Physics.DYN, ctrl_freq=ctrl_freq) for _ in range(n)]
        for e in envs:
            e.reset()
        a = np.zeros((1, 4), np.float32)
        for e in envs:
            e.step(a)
        t0 = time.perf_counter()
        for _ in range(steps):
            for e in envs:
                e.step(a)
        return (time.perf_counter() - t0) / (steps * n)


def run(steps: int, warmup: int, ctrl_freq: int = 30, workers: int | None = None, envs_per_worker: int | None = None,
        budget_s: float = 30.0, seed: int = 0):
    """`steps` timed vec-steps (after `warmup`) of workers x envs_per_worker reference envs.  Returns a dict with
    value = drone-substeps/s (1 drone per env, S = 240 // ctrl_freq substeps per step)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (python oracle/stage_reference.py, needs /root/reference)")
    workers = workers or os.cpu_count() or 1
    S = 240 // ctrl_freq
    if envs_per_worker is None:
        per_step = calibrate(ctrl_freq)
        envs_per_worker = int(max(1, min(512, budget_s / (per_step * max(1, steps + warmup) * 1.3))))
    ctx = mp.get_context("fork")
    start = ctx.Barrier(workers + 1)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(w, envs_per_worker, warmup, steps, ctrl_freq, seed, start, q), daemon=True)
             for w in range(workers)]
    for p in procs:
        p.start()
    try:
        start.wait(timeout=600)
        t0 = time.perf_counter()
        res = [q.get(timeout=1200) for _ in procs]
        wall = time.perf_counter() - t0
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    errs = [r[3] for r in res if r[3]]
    if errs:
        raise RuntimeError("reference worker failed: " + errs[0])
    slowest = max(r[1] for r in res)
    env_steps = workers * envs_per_worker * steps
    return dict(value=env_steps * S / slowest, env_steps=env_steps, seconds=slowest, wall=wall, workers=workers,
                envs_per_worker=envs_per_worker, steps=steps, S=S, resets=sum(r[2] for r in res),
                per_core=env_steps * S / slowest / workers)


if __name__ == "__main__":
    import json
    r = run(steps=int(sys.argv[1]) if len(sys.argv) > 1 else 20, warmup=3)
    print(json.dumps(r))
