#!/usr/bin/env python
"""Reads gpurun_out/prof.ncu-rep (ncu --set full of the step kernel under `bench.py --no-graph`) and writes
profiles/traffic.json: DRAM bytes per launch of the dominant kernel, which bench.py reports as roofline.traffic."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof.ncu-rep")
big = sys.argv[2] if len(sys.argv) > 2 else None          # optional: capture of the same kernel at >= 1 M envs (multi-wave)


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return r, r[0], r[1]


rows, hdr, units = load(rep)


def col(name):
    i = hdr.index(name)
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[i], 1.0)
    return [float(r[i].replace(",", "")) * scale for r in rows[2:]]


rd, wr, dur = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
n = len(rd)
res = {"kernel": rows[2][hdr.index("Kernel Name")], "launches_profiled": n,
       "dram_bytes_read_per_launch": sum(rd) / n, "dram_bytes_write_per_launch": sum(wr) / n,
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n, "gpu_time_us_under_ncu": sum(dur) / n,
       "note": "ncu --set full --clock-control none; per-launch times are cold-cache and serialised; with the default "
               "cache control L2 is flushed before each replay, so writes still resident in L2 at kernel end are not counted",
       "source": os.path.relpath(rep, ROOT)}
if big:
    # a multi-wave launch evicts its own writes while it runs, so its DRAM counters see reads AND writes: the steady-state
    # bytes per env-step, which the rotating env sets of bench.py pay at every size
    rows, hdr, units = load(big)
    rd2, wr2, grid = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("launch__grid_size")
    envs = grid[0] * 64
    res["steady_state_bytes_per_env_step"] = (sum(rd2) + sum(wr2)) / len(rd2) / envs
    res["steady_state_source"] = os.path.relpath(big, ROOT) + f" ({int(envs)} envs per launch)"
json.dump(res, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(res))
