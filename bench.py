#!/usr/bin/env python
"""bench.py — drone-substeps/sec of the fused DYN step kernel (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the oracle port on the host cores

Workload (BASELINE.json configs[1]): HoverAviary single-drone PPO-rollout shape, 65,536 parallel envs per GPU,
Physics.DYN, ActionType.RPM, KIN observation (72 floats), FP32, 240 Hz sim / 30 Hz ctrl (8 substeps per step),
uniform random float32 actions, SB3-style auto-reset on.  One "step" = one env.step() of all 65,536 envs
= one launch of the fused kernel.  The 42 MB per-step working set fits the 126 MB L2, so the bench rotates
over `--sets` independent env sets (8 x ~46 MB > L2): every step touches data last used 8 steps ago.
Steps are replayed from one CUDA graph (2*sets kernel nodes: the observation ping-pong has period 2) so the
host launch rate does not bound a ~10 us kernel.

value   = E*N*S*K / device time (CUDA events, max over ranks), inputs resident in HBM
e2e     = the same metric through HoverAviary.step(numpy) : pinned host action -> H2D, kernel, D2H obs/reward/flags
roofline= algorithmic bytes (646 B per env-step) * E / kernel time, against MEASURED_PEAKS.json hbm_gbs
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_ENV_STEP = {30: 646, 48: 934}      # SURVEY §8d, FP32, HoverAviary RPM KIN
# FP64 mode: the 13-value state and the reward double; actions, ring and observation stay float32
ALGO_BYTES_PER_ENV_STEP_F64 = {30: 646 + 2 * 52 + 4, 48: 934 + 2 * 52 + 4}
METRIC = "drone-substeps/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--ctrl-freq", type=int, default=30)
    ap.add_argument("--sets", type=int, default=8, help="independent env sets rotated to defeat L2 residency")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--tpb", type=int, default=0)
    ap.add_argument("--streams", type=int, default=1,
                    help="async env pools: env set j always steps on stream j %% STREAMS, so independent sets overlap (default 1 = "
                         "every step ordered on one stream, the headline mode)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-async-extra", action="store_true", help="skip the informational multi-stream measurement")
    return ap.parse_args()


def workload_name(a):
    return (f"HoverAviary DYN RPM KIN {a.precision} {a.envs} envs/GPU x 1 drone, 240/{a.ctrl_freq} Hz "
            f"(S={240 // a.ctrl_freq}), U(-1,1) float32 actions, auto-reset")


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/gpd_oracle.c) on the host cores.  The reference itself is pure Python and
# cannot travel to the GPU box (no /root/reference there), so kind = "port".
def cpu_run(a, seconds=None, steps=None, warmup=2, budget_s=120.0):
    """The oracle port on all host cores.  With a fixed number of `steps` the per-step sample (number of envs, at most
    the bench's own) is sized from a calibration step so that the whole run stays within `budget_s`: the metric is
    per drone-substep, and the CPU cost is linear in the number of envs."""
    from gpd_b200.params import load_drone_params
    from gpd_b200.utils.enums import DroneModel
    from oracle import oracle as orc
    E = a.envs
    threads = orc.max_threads()
    sim = orc.OracleSim(load_drone_params(DroneModel.CF2X), E, ctrl_freq=a.ctrl_freq)
    rng = np.random.default_rng(0)
    pool = [rng.uniform(-1, 1, size=(E, 1, 4)).astype(np.float32) for _ in range(4)]
    S = 240 // a.ctrl_freq

    def one(k):
        _, _, te, tr = sim.step(pool[k % 4], nthreads=threads)
        done = (te | tr)
        if done.any():
            sim.reset(done)
        return sim
    for k in range(warmup):
        one(k)
    if steps is not None:
        # size the per-step sample from two calibration steps of the real loop; never below 8192 envs, where the
        # per-step thread fork/join would start to dominate and under-state the CPU
        t0 = time.perf_counter()
        one(0); one(1)
        per_step = (time.perf_counter() - t0) / 2
        if per_step * steps > budget_s and E > 8192:
            E = max(8192, int(E * budget_s / (per_step * steps)))
            sim = orc.OracleSim(load_drone_params(DroneModel.CF2X), E, ctrl_freq=a.ctrl_freq)
            pool = [p[:E].copy() for p in pool]
            for k in range(warmup):
                one(k)
    t0 = time.perf_counter()
    n = 0
    while True:
        one(n)
        n += 1
        el = time.perf_counter() - t0
        if steps is not None and n >= steps:
            break
        if steps is None and el >= seconds:
            break
    el = time.perf_counter() - t0
    return dict(value=E * S * n / el, steps=n, seconds=el, cores=threads, E=E, S=S)


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_run(a, steps=max(1, a.steps), warmup=max(1, a.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "drone-substeps/s", "n_gpus": a.gpus,
        "steps": r["steps"], "warmup": a.warmup, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": r["value"], "unit": "drone-substeps/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{r['steps']} env.step() of {r['E']} envs (FP64 C oracle port, pthreads, auto-reset)"},
        "e2e": {"value": r["value"], "unit": "drone-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi style clock/throttle sampling (NVML) while the GPU is busy."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:        # every sample between start() and stop() is taken while the timed regions run
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append(mhz)
                for k, bit in names.items():
                    if rs & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_from_profile():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def b200_arm(a):
    import torch
    import torch.distributed as dist

    import gpd_b200  # noqa: F401
    from gpd_b200.envs import HoverAviary
    from gpd_b200.utils.enums import ActionType, ObservationType, Physics

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    E, S, nsets = a.envs, 240 // a.ctrl_freq, a.sets

    envs = [HoverAviary(physics=Physics.DYN, ctrl_freq=a.ctrl_freq, obs=ObservationType.KIN, act=ActionType.RPM,
                        num_envs=E, device=local, precision=a.precision, auto_reset=True, threads_per_block=a.tpb)
            for _ in range(nsets)]
    g = torch.Generator(device=dev)
    g.manual_seed(rank)
    acts = [(torch.rand((E, 1, 4), generator=g, device=dev) * 2 - 1) for _ in range(2 * nsets)]
    for e in envs:
        e.reset()
    period = 2 * nsets

    nstreams = max(1, min(a.streams, nsets))
    pool = [torch.cuda.Stream(device=dev) for _ in range(nstreams - 1)]

    def run_steps(n, nstreams=nstreams, pool=pool):
        if nstreams == 1:
            for k in range(n):
                envs[k % nsets]._sim.step(acts[k])
            return
        # fork: every pool stream waits for the caller's stream; set j always runs on stream j % nstreams (its own steps stay
        # ordered); join: the caller's stream waits for every pool stream.  All of it is capturable.
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        for st in pool:
            st.wait_event(fork)
        for k in range(n):
            j = (k % nsets) % nstreams
            with torch.cuda.stream(main if j == 0 else pool[j - 1]):
                envs[k % nsets]._sim.step(acts[k])
        for st in pool:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    def cycle():
        run_steps(period)

    # warm-up (also touches every buffer)
    wu = max(3, a.warmup)
    for _ in range((wu + period - 1) // period):
        cycle()
    torch.cuda.synchronize()
    K = max(1, a.steps)
    reps, tail = divmod(K, period)          # EXACTLY K steps: `reps` replays of the 16-step graph + one tail graph

    def run_tail():
        run_steps(tail)
    graph = tail_graph = None
    graph_error = None

    def capture():
        side = torch.cuda.Stream()
        g_main, g_tail = torch.cuda.CUDAGraph(), None
        with torch.cuda.stream(side):
            with torch.cuda.graph(g_main, stream=side):
                cycle()
            if tail:
                saved = [(e._sim._cur, e._sim._have_prev) for e in envs]
                g_tail = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_tail, stream=side):
                    run_tail()
                for e, (c, h) in zip(envs, saved):      # capture only records: restore the host-side ping-pong phase
                    e._sim._cur, e._sim._have_prev = c, h
        torch.cuda.synchronize()
        g_main.replay()
        torch.cuda.synchronize()
        return g_main, g_tail

    if not a.no_graph:
        try:
            graph, tail_graph = capture()
        except Exception as ex:          # never lose the measurement to a capture problem: fall back to direct launches
            graph = tail_graph = None
            graph_error = repr(ex)[:200]
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if graph is not None:           # bring the clocks to their loaded state before sampling starts
        for _ in range(max(1, min(reps // 4, 500))):
            graph.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        if graph is not None:
            graph.replay()
        else:
            cycle()
    if tail:
        if tail_graph is not None:
            tail_graph.replay()
        else:
            run_tail()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    value = world * E * 1 * S * K / (ms * 1e-3)

    # ---- informational: the same K-step job with the env sets as async pools on their own streams (never the headline) ----
    async_info = None
    if nstreams == 1 and graph is not None and nsets >= 2 and not a.no_async_extra:
        try:
            ns2 = min(8, nsets)
            pool2 = [torch.cuda.Stream(device=dev) for _ in range(ns2 - 1)]
            side = torch.cuda.Stream()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g2, stream=side):
                    run_steps(period, ns2, pool2)
            torch.cuda.synchronize()
            reps2 = max(1, min(reps, 250))
            for _ in range(20):
                g2.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps2):
                g2.replay()
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / (reps2 * period)
            async_info = {"streams": ns2, "ms_per_step": ms2, "steps": reps2 * period,
                          "value_this_rank": E * S / (ms2 * 1e-3),
                          "note": "env set j steps on stream j % streams (its own steps stay ordered, independent sets overlap): "
                                  "what a trainer with several env pools gets (pool.py); informational, not the headline"}
        except Exception as ex:
            async_info = {"error": repr(ex)[:200]}
            torch.cuda.synchronize()

    # ---- informational: one env set alone (46 MB working set: L2-resident; SURVEY 8d asks for flushed AND unflushed) ----
    l2_info = None
    if graph is not None and not a.no_async_extra:
        try:
            side = torch.cuda.Stream()
            g3 = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g3, stream=side):
                    for k in range(period):
                        envs[0]._sim.step(acts[k % 2])
            torch.cuda.synchronize()
            reps3 = max(1, min(reps, 250))
            for _ in range(20):
                g3.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps3):
                g3.replay()
            e1.record()
            torch.cuda.synchronize()
            l2_info = {"ms_per_step": e0.elapsed_time(e1) / (reps3 * period), "steps": reps3 * period,
                       "note": "ONE env set stepped back to back (working set < 126 MB L2, no rotation): what a single-set "
                               "rollout sees; not an HBM number, informational"}
        except Exception as ex:
            l2_info = {"error": repr(ex)[:200]}
            torch.cuda.synchronize()

    # ---- episode statistics: the only collective (NCCL all-reduce of 6 sums + min/max), outside the step path ----
    stats = np.zeros(8)
    for e in envs:
        s = e._sim.episode_stats()
        stats[[0, 1, 2, 3, 6, 7]] += s[[0, 1, 2, 3, 6, 7]]
    if world > 1:
        t = torch.tensor(stats, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        stats = t.cpu().numpy()

    # ---- end to end through the public API with host buffers (rank-local; aggregate = sum over ranks) ----
    env = envs[0]
    pinned_act = [torch.empty((E, 1, 4), dtype=torch.float32).pin_memory() for _ in range(4)]
    rng = np.random.default_rng(rank)
    for p in pinned_act:
        p.copy_(torch.from_numpy(rng.uniform(-1, 1, size=(E, 1, 4)).astype(np.float32)))
    np_act = [p.numpy() for p in pinned_act]
    for k in range(3):
        env.step(np_act[k % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for k in range(a.e2e_steps):
        obs, rew, term, trunc, _ = env.step(np_act[k % 4])
        chk += float(rew[0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * E * S * a.e2e_steps / e2e_s
    W = env._sim.W
    h2d = E * 4 * 4
    d2h = E * (W * 4 + 4 + 1 + 1)      # obs + reward + terminated + truncated, one packed copy
    clocks = sampler.stop()

    if rank == 0:
        peak, peak_src = peaks()
        per_launch_ms = ms / K
        algo = (ALGO_BYTES_PER_ENV_STEP_F64 if a.precision == "f64" else ALGO_BYTES_PER_ENV_STEP).get(a.ctrl_freq, None)
        achieved = (algo * E / (per_launch_ms * 1e-3) / 1e9) if algo else None
        tr = traffic_from_profile()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": (tr or {}).get("dram_bytes_per_launch"), "traffic_note": (tr or {}).get("note"),
                "traffic_steady_state": ((tr or {}).get("steady_state_bytes_per_env_step") or 0) * E or None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo * E if algo else None, "kernel": "gpd::step_kernel<%s,LEAN,N=1,VEC>" % ("double" if a.precision == "f64" else "float"),
                "kernel_ms_per_launch": per_launch_ms}
        if nstreams > 1:
            roof["note"] = (f"{nstreams} streams: launches of independent env sets overlap, so kernel_ms_per_launch is the "
                            "throughput-equivalent time per launch, not one launch's duration")
        if async_info and "ms_per_step" in async_info and algo:
            async_info["frac_of_hbm_peak"] = algo * E / (async_info["ms_per_step"] * 1e-3) / 1e9 / peak
        cpu = None
        if not a.no_cpu:
            r = cpu_run(a, seconds=a.cpu_seconds)
            cpu = {"value": r["value"], "unit": "drone-substeps/s", "cores": r["cores"], "kind": "port",
                   "sample": f"{r['steps']} env.step() of {r['E']} envs in {r['seconds']:.1f} s (FP64 C oracle port, pthreads, auto-reset)"}
        line = {
            "metric": METRIC, "value": value, "unit": "drone-substeps/s", "n_gpus": world, "steps": K, "warmup": wu,
            "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.precision, "data": "synthetic",
            "config": {"workload": workload_name(a), "envs_per_gpu": E, "substeps_per_step": S,
                       "l2": f"rotating {nsets} env sets per GPU (~{nsets * E * 700 / 1e6:.0f} MB touched per cycle > 126 MB L2)",
                       "launch": ("CUDA graph of %d step kernels x %d replays + %d-step tail graph" % (period, reps, tail))
                       if graph is not None else ("direct launches" + (f" (graph capture failed: {graph_error})" if graph_error else "")),
                       "streams": nstreams,
                       "parallelism": f"env-sharded x{world}, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "drone-substeps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": a.e2e_steps, "api": "HoverAviary.step(numpy) -> gpd_step_host (pinned host buffers)"},
            "gpu_launches": K,
            "roofline": roof,
            "async_pools": async_info,
            "l2_resident": l2_info,
            "cpu_baseline": cpu,
            "episode_stats": {"episodes": stats[0], "mean_return": stats[1] / max(stats[0], 1),
                              "mean_length": stats[2] / max(stats[0], 1), "env_steps": stats[6]},
        }
        print(json.dumps(line), flush=True)
    for e in envs:
        e.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)


if __name__ == "__main__":
    main()
