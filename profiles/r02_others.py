#!/usr/bin/env python
"""C3 / C4 / C5 (bench.py other_configs) under a scheduling variant given by the environment: prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

timer = bench.Timer(torch, dist, 1, torch.device("cuda", 0))
o = bench.other_configs(torch, timer, 1, 0, 0, bench.peaks()[0])
print(json.dumps({"label": os.environ.get("LABEL", ""), **{k: (round(v["us_per_step"], 2) if "us_per_step" in v else v) for k, v in o.items()}}))
