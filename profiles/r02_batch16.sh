#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 9 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b16.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b16.jsonl
}
run "1M tpb=64" -- --steps 48 --warmup 5 --envs 1048576 --sets 2 --tpb 64
run "1M tpb=128" -- --steps 48 --warmup 5 --envs 1048576 --sets 2 --tpb 128
run "1M tpb=96" -- --steps 48 --warmup 5 --envs 1048576 --sets 2 --tpb 96
run "262144 tpb=64" -- --steps 96 --warmup 5 --envs 262144 --sets 4 --tpb 64
run "262144 tpb=128" -- --steps 96 --warmup 5 --envs 262144 --sets 4 --tpb 128
run "524288 tpb=64" -- --steps 64 --warmup 5 --envs 524288 --sets 2 --tpb 64
run "524288 tpb=128" -- --steps 64 --warmup 5 --envs 524288 --sets 2 --tpb 128
run "32768 tpb=64" -- --steps 200 --warmup 5 --envs 32768 --sets 16 --tpb 64
run "32768 tpb=128" -- --steps 200 --warmup 5 --envs 32768 --sets 16 --tpb 128
run "49152 tpb=64" -- --steps 200 --warmup 5 --envs 49152 --sets 12 --tpb 64
run "49152 tpb=128" -- --steps 200 --warmup 5 --envs 49152 --sets 12 --tpb 128
run "65536 default K=20" -- --steps 20 --warmup 5
cat > /tmp/c5.py <<'PY'
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
import bench
from gpd_b200.envs import HoverAviary
from gpd_b200.utils.enums import ActionType, DroneModel
timer = bench.Timer(torch, dist, 1, torch.device("cuda", 0))
E = 2097152
g = torch.Generator(device="cuda"); g.manual_seed(0)
for tpb in (0, 64, 96, 128):
    r = bench.measure_config(torch, timer, lambda: HoverAviary(num_envs=E, drone_model=DroneModel.CF2P, ctrl_freq=48, act=ActionType.PID, precision="f32", auto_reset=True, threads_per_block=tpb),
                             lambda env, k: torch.rand((E, 1, 3), generator=g, device="cuda") * 2 - 1, nsets=1, steps=6)
    print(json.dumps({"c5 tpb": tpb, "us": round(r["us_per_step"], 1)}))
PY
python /tmp/c5.py 2>/dev/null | tail -4
