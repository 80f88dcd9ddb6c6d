#!/usr/bin/env python
"""Sorted registers / stack / spill summary of every kernel in the shipped build, from csrc/obj/ptxas_*.log
(`-Xptxas -v`, written by csrc/build.sh).  Usage: python profiles/ptxas_summary.py > profiles/rNN/ptxas_summary.txt"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
logs = sorted(glob.glob(os.path.join(ROOT, "gym-pybullet-drones-routing_b200", "csrc", "obj", "ptxas_*.log")))
rows = {}
for p in logs:
    name = None
    stack = spill_st = spill_ld = 0
    for line in open(p):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
            stack = spill_st = spill_ld = 0
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            stack, spill_st, spill_ld = map(int, m.groups())
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            rows[name] = (int(m.group(1)), stack, spill_st, spill_ld)
            name = None
names = list(rows)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
out = []
for mangled, d in zip(names, dem):
    d = re.sub(r"\(.*", "", d).replace("void gpd::", "").replace("void ", "")
    r = rows[mangled]
    out.append(f"{d:<70} regs={r[0]:3d} stack={r[1]:4d} spill_st={r[2]:4d} spill_ld={r[3]:4d}")
print("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -v (csrc/build.sh): registers / stack / spill bytes per kernel")
print("# step_kernel<Real, KIND, MULTI, VEC>: KIND 0 = force models, 1 = lean, 2 = DSLPID in the loop")
for l in sorted(out):
    print(l)
