#!/usr/bin/env python
"""Throughput of the other BASELINE.json shapes (parity-test configs, not bench lines), one GPU, CUDA-graph replay,
rotating env sets sized to exceed the 126 MB L2.  Algorithmic bytes per env-ctrl-step from SURVEY §8d.

    python profiles/configs.py [name ...] > profiles/rNN/configs.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import CtrlAviary, HoverAviary, MultiHoverAviary  # noqa: E402
from gpd_b200.utils.enums import ActionType, DroneModel, Physics  # noqa: E402

PEAK = 6553.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def measure(make_env, make_action, nsets, reps, trials=3):
    envs = [make_env() for _ in range(nsets)]
    acts = [make_action(envs[0], k) for k in range(2 * nsets)]
    for e in envs:
        e.reset()
        e._sim.set_step_chaining(True)      # the action buffers are generated before the first step is enqueued (gpd.h)
    period = 2 * nsets

    def cycle():
        for k in range(period):
            envs[k % nsets]._sim.step(acts[k])
    for _ in range(2):
        cycle()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            cycle()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(trials):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (reps * period) * 1e3)
    sim = envs[0]._sim
    info = dict(E=sim.E, N=sim.N, S=sim.S, W=sim.W, us_per_step=best,
                drone_substeps_per_s=sim.E * sim.N * sim.S / (best * 1e-6))
    if getattr(envs[0], "AUTO_RESET", False) or getattr(sim, "auto_reset", False):
        st = sim.episode_stats()         # include/gpd.h: episodes, sum return, sum length, sum return^2, min, max, env-steps, terminated
        if st[6] > 0 and st[0] > 0:      # (each env set replays 2 action batches, so episodes are shorter than under fresh noise)
            info.update(resets_per_env_step=st[0] / st[6], mean_episode_ctrl_steps=st[2] / st[0], mean_episode_return=st[1] / st[0])
    for e in envs:
        e.close()
    del envs, acts, gr
    torch.cuda.empty_cache()
    return info


def rand_act(shape, seed, lo=-1.0, hi=1.0, dtype=torch.float32):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return (torch.rand(shape, generator=g, device="cuda", dtype=torch.float32) * (hi - lo) + lo).to(dtype)


CONFIGS = {}


def config(name):
    def deco(f):
        CONFIGS[name] = f
        return f
    return deco


@config("c2_hover_30hz_f32")
def c2_30():
    E = 65536
    r = measure(lambda: HoverAviary(num_envs=E, ctrl_freq=30, precision="f32", auto_reset=True),
                lambda env, k: rand_act((E, 1, 4), k), nsets=8, reps=40)
    r.update(algo_bytes_per_env_step=646, workload="HoverAviary 65,536 envs RPM KIN FP32 240/30")
    return r


def _c2_stream(label, make_action, **env_kw):
    E = 65536
    r = measure(lambda: HoverAviary(num_envs=E, ctrl_freq=30, precision="f32", auto_reset=True, **env_kw),
                make_action, nsets=8, reps=40)
    r.update(algo_bytes_per_env_step=646, workload="HoverAviary 65,536 envs RPM KIN FP32 240/30, " + label)
    return r


@config("c2_hover_30hz_f32_random_init")
def c2_random_init():
    """xyz ~ U([-1,1]^2 x [0.2,1.5]), rpy ~ U(-0.2,0.2)^3, uniform actions."""
    E = 65536
    rng = np.random.default_rng(0)
    xyz = np.concatenate([rng.uniform(-1, 1, size=(E, 1, 2)), rng.uniform(0.2, 1.5, size=(E, 1, 1))], axis=-1)
    rpy = rng.uniform(-0.2, 0.2, size=(E, 1, 3))
    return _c2_stream("randomised initial poses", lambda env, k: rand_act((E, 1, 4), k), initial_xyzs=xyz, initial_rpys=rpy)


@config("c2_hover_30hz_f32_near_hover")
def c2_near_hover():
    """a ~ N(0, 0.05^2) (longer episodes, still tilt-truncated)."""
    E = 65536

    def act(env, k):
        g = torch.Generator(device="cuda")
        g.manual_seed(k)
        return torch.randn((E, 1, 4), generator=g, device="cuda") * 0.05
    return _c2_stream("near-hover actions N(0, 0.05^2)", act)


@config("c2_hover_30hz_f32_symmetric")
def c2_symmetric():
    """all four motors equal -- the only stream that reaches the 8 s time limit."""
    E = 65536
    return _c2_stream("symmetric actions (four equal motors)", lambda env, k: rand_act((E, 1, 1), k, -0.05, 0.05).expand(E, 1, 4).contiguous())


@config("c2_hover_48hz_f32")
def c2_48():
    E = 65536
    r = measure(lambda: HoverAviary(num_envs=E, ctrl_freq=48, precision="f32", auto_reset=True),
                lambda env, k: rand_act((E, 1, 4), k), nsets=6, reps=40)
    r.update(algo_bytes_per_env_step=934, workload="HoverAviary 65,536 envs RPM KIN FP32 240/48")
    return r


@config("c3_multihover2_gnd_drag_f64")
def c3():
    E = 32768
    r = measure(lambda: MultiHoverAviary(num_envs=E, num_drones=2, physics=Physics.DYN_GND_DRAG, ctrl_freq=30,
                                         precision="f64", auto_reset=True),
                lambda env, k: rand_act((E, 2, 4), k), nsets=6, reps=30)
    # 1,292 B per drone-ctrl-step in FP64 accounting (obs stays float32 here: state/aux in double)
    r.update(algo_bytes_per_env_step=2 * 1292, workload="MultiHoverAviary 32,768 envs x 2 drones DYN+GND+DRAG FP64 240/30")
    return r


@config("c3_multihover2_gnd_drag_f32")
def c3_f32():
    E = 32768
    r = measure(lambda: MultiHoverAviary(num_envs=E, num_drones=2, physics=Physics.DYN_GND_DRAG, ctrl_freq=30,
                                         precision="f32", auto_reset=True),
                lambda env, k: rand_act((E, 2, 4), k), nsets=8, reps=30)
    r.update(algo_bytes_per_env_step=2 * 646, workload="MultiHoverAviary 32,768 envs x 2 drones DYN+GND+DRAG FP32 240/30")
    return r


def _c4(precision):
    E, N = 4096, 64
    rng = np.random.default_rng(1)
    xyz = np.concatenate([rng.uniform(-2, 2, size=(E, N, 2)), rng.uniform(0.2, 3, size=(E, N, 1))], axis=-1)
    dt = torch.float64 if precision == "f64" else torch.float32

    def mk():
        return CtrlAviary(num_envs=E, num_drones=N, physics=Physics.DYN_DW, pyb_freq=240, ctrl_freq=48,
                          initial_xyzs=xyz, precision=precision)

    def act(env, k):
        return (env.HOVER_RPM * (1 + 0.02 * rand_act((E, N, 4), k))).to(dt)
    r = measure(mk, act, nsets=4, reps=10)
    r.update(algo_bytes_per_env_step=N * 200 * (2 if precision == "f64" else 1),
             workload=f"CtrlAviary 4,096 envs x 64 drones DYN+DW {precision} 240/48 (O(N^2) downwash)",
             pair_evals_per_s=E * N * N * r["S"] / (r["us_per_step"] * 1e-6))
    return r


@config("c4_ctrl64_dw_f32")
def c4_f32():
    return _c4("f32")


@config("c4_ctrl64_dw_f64")
def c4_f64():
    return _c4("f64")


@config("c5_hover_pid_48hz_f32")
def c5():
    E = 2097152
    r = measure(lambda: HoverAviary(num_envs=E, drone_model=DroneModel.CF2P, ctrl_freq=48, act=ActionType.PID,
                                    precision="f32", auto_reset=True),
                lambda env, k: rand_act((E, 1, 3), k), nsets=1, reps=6)
    r.update(algo_bytes_per_env_step=814, workload="HoverAviary 2,097,152 envs ActionType.PID (DSLPID in-loop) FP32 240/48")
    return r


@config("c5_ctrl_rollout_pid_f32")
def c5_rollout():
    """pid.py-style loop with the controller in the kernel: gpd_rollout_pid, 48 ctrl steps per launch."""
    from gpd_b200.params import default_pid_params, load_drone_params
    from gpd_b200.sim import BatchedSim
    E, steps = 2097152, 48
    dp = load_drone_params(DroneModel.CF2P)
    sim = BatchedSim(dp, E, 1, env_kind="ctrl", action_type="ctrl_rpm", pyb_freq=240, ctrl_freq=48, precision="f32",
                     pid=default_pid_params(DroneModel.CF2P))
    n_wp = 480
    ang = np.arange(n_wp) / n_wp * 2 * np.pi + np.pi / 2
    wps = torch.tensor(np.stack([.3 * np.cos(ang), .3 * np.sin(ang) - .3, np.zeros(n_wp)], axis=1), dtype=torch.float32,
                       device="cuda")
    wp = torch.randint(0, n_wp, (E, 1), dtype=torch.int32, device="cuda")
    act = torch.zeros((E, 1, 4), dtype=torch.float32, device="cuda")
    sim.rollout_pid(steps, wps, wp, act)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.rollout_pid(steps, wps, wp, act)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    r = dict(E=E, N=1, S=5, us_per_launch=best, ctrl_steps_per_launch=steps,
             drone_substeps_per_s=E * 5 * steps / (best * 1e-6),
             workload="CtrlAviary 2,097,152 envs, DSLPID + circle waypoints in-kernel, 48 ctrl steps per launch, FP32 240/48",
             bound="FP32/SFU issue (state stays in registers: 2 x 100 B of HBM traffic per drone per launch)")
    sim.close()
    return r


if __name__ == "__main__":
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        r = CONFIGS[n]()
        r["config"] = n
        if "algo_bytes_per_env_step" in r and "us_per_step" in r:
            gbs = r["algo_bytes_per_env_step"] * r["E"] / (r["us_per_step"] * 1e-6) / 1e9
            r["achieved_gbs"] = gbs
            r["frac_of_hbm_peak"] = gbs / PEAK
        print(json.dumps(r), flush=True)
