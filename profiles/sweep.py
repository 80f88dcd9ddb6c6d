#!/usr/bin/env python
"""In-process tuning sweep for the step kernel (same GPU, same clocks for every variant).
Each variant = env-var overrides read by gpd_create + bench-style rotating env sets replayed from a CUDA graph."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpd_b200  # noqa: E402,F401
from gpd_b200.envs import HoverAviary  # noqa: E402
from gpd_b200.utils.enums import Physics  # noqa: E402


def run(E=65536, nsets=8, tpb=0, env=None, auto_reset=True, reps=60, trials=3, ctrl_freq=30, precision="f32", graph=True, pyb_freq=240, streams=1):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = str(v)
    envs = [HoverAviary(physics=Physics.DYN, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, num_envs=E, precision=precision, auto_reset=auto_reset,
                        threads_per_block=tpb) for _ in range(nsets)]
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    acts = [(torch.rand((E, 1, 4), generator=g, device="cuda") * 2 - 1) for _ in range(2 * nsets)]
    for e in envs:
        e.reset()
    period = 2 * nsets

    extra = [torch.cuda.Stream() for _ in range(streams - 1)]

    def cycle():
        if streams == 1:
            for k in range(period):
                envs[k % nsets]._sim.step(acts[k])
            return
        # independent env sets round-robin over `streams` streams (fork/join with events; capturable)
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        for st in extra:
            st.wait_event(fork)
        for k in range(period):
            st = main if k % streams == 0 else extra[k % streams - 1]
            with torch.cuda.stream(st):
                envs[k % nsets]._sim.step(acts[k])
        for st in extra:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    gr = None
    if graph:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                cycle()
        torch.cuda.synchronize()
        gr.replay()
    torch.cuda.synchronize()
    best = []
    for _ in range(trials):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay() if gr is not None else cycle()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / (reps * period) * 1e3)
    for e in envs:
        e.close()
    del envs, acts, gr
    torch.cuda.empty_cache()
    return best


def empty_graph_latency(n=16, reps=200):
    """per-node latency of a dependent chain of trivial kernels replayed from a CUDA graph"""
    x = torch.zeros(32, device="cuda")
    for _ in range(3):
        x.add_(1)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                x.add_(1)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n) * 1e3


if __name__ == "__main__":
    variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else [{}]
    if len(sys.argv) > 2 and sys.argv[2] == "empty":
        print(json.dumps({"empty_graph_node_us": round(empty_graph_latency(), 3)}), flush=True)
    for v in variants:
        kw = dict(v)
        envv = kw.pop("env", None)
        us = run(env=envv, **kw)
        E = kw.get("E", 65536)
        f = kw.get("ctrl_freq", 30)
        algo = {30: 646, 48: 934}[f]
        print(json.dumps({"variant": v, "us_per_step": [round(x, 3) for x in us],
                          "frac_of_6553": round(algo * E / (min(us) * 1e-6) / 6553e9, 4)}), flush=True)
