#!/usr/bin/env bash
# (1) how much of the step is the per-tile sequencing itself: UNSAFE experiment switches (results are discarded)
# (2) ncu --set full of the C3 and C5 kernels (other_configs)
set -u
O=gpurun_out/r02
mkdir -p $O
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b6.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b6.jsonl
}
run "safe K=200" -- --steps 200 --warmup 5
run "UNSAFE publish without waiting for the stores K=200" GPD_DEBUG_UNSAFE=1 -- --steps 200 --warmup 5
run "UNSAFE no claim/publish at all K=200" GPD_DEBUG_UNSAFE=3 -- --steps 200 --warmup 5
run "safe K=20" -- --steps 20 --warmup 5
run "UNSAFE publish without waiting K=20" GPD_DEBUG_UNSAFE=1 -- --steps 20 --warmup 5
run "UNSAFE no claim/publish K=20" GPD_DEBUG_UNSAFE=3 -- --steps 20 --warmup 5
cat > /tmp/one_cfg.py <<'PY'
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import gpd_b200
from gpd_b200.envs import HoverAviary, MultiHoverAviary
from gpd_b200.utils.enums import ActionType, DroneModel, Physics
which = sys.argv[1]
g = torch.Generator(device="cuda"); g.manual_seed(0)
if which == "c3":
    E = 32768
    envs = [MultiHoverAviary(num_envs=E, num_drones=2, physics=Physics.DYN_GND_DRAG, ctrl_freq=30, precision="f64", auto_reset=True) for _ in range(6)]
    acts = [torch.rand((E, 2, 4), generator=g, device="cuda") * 2 - 1 for _ in range(4)]
else:
    E = 2097152
    envs = [HoverAviary(num_envs=E, drone_model=DroneModel.CF2P, ctrl_freq=48, act=ActionType.PID, precision="f32", auto_reset=True)]
    acts = [torch.rand((E, 1, 3), generator=g, device="cuda") * 2 - 1 for _ in range(4)]
for e in envs:
    e.reset(); e._sim.set_step_chaining(True)
for k in range(24):
    envs[k % len(envs)]._sim.step(acts[k % 4])
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 3 -f -o $O/prof_c3 python /tmp/one_cfg.py c3 > $O/ncu_c3.log 2>&1; echo "c3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 2 -f -o $O/prof_c5 python /tmp/one_cfg.py c5 > $O/ncu_c5.log 2>&1; echo "c5 rc=$?"
ls -la $O/*.ncu-rep
