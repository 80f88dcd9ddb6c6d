#!/usr/bin/env python
"""Block-size sweep of the multi-drone shapes (same process, same GPU): python profiles/tpb_sweep.py [f32|f64] ..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import configs  # noqa: E402
from configs import MultiHoverAviary, Physics, measure, rand_act  # noqa: E402

if __name__ == "__main__":
    E = 32768
    for prec in (sys.argv[1:] or ["f32"]):
        for tpb in (0, 64, 96, 128, 160, 192, 224, 256):
            r = measure(lambda: MultiHoverAviary(num_envs=E, num_drones=2, physics=Physics.DYN_GND_DRAG, ctrl_freq=30,
                                                 precision=prec, auto_reset=True, threads_per_block=tpb),
                        lambda env, k: rand_act((E, 2, 4), k), nsets=8 if prec == "f32" else 6, reps=20)
            print(json.dumps(dict(precision=prec, tpb=tpb, us_per_step=round(r["us_per_step"], 2))), flush=True)
