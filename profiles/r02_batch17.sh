#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_b17.log
tail -5 $O/pytest_b17.log
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_b17_full.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_b17_full.json"))
print("headline", d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"])
print("async", d["async_pools"]["ms_per_step"], "l2_resident", d["l2_resident"]["ms_per_step"])
for k, v in d["other_configs"].items():
    print(k, v.get("us_per_step"), v.get("roofline", {}).get("frac"), v.get("error"))
PY
python bench.py --steps 48 --warmup 5 --envs 1048576 --sets 2 --no-extra --no-cpu --e2e-steps 2 --trials 9 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1M', d['ms_per_step'], d['roofline']['frac'])"
CMD3="python bench.py --steps 6 --warmup 3 --no-cpu --no-extra --no-graph --e2e-steps 2 --envs 1048576 --sets 2"
ncu --set full --clock-control none -k regex:step_kernel -s 10 -c 2 -f -o $O/final/prof_bulk_1M $CMD3 > $O/final/ncu_full_1M.log 2>&1
echo "full 1M rc=$?"
cat > /tmp/one_cfg.py <<'PY'
import os, sys
import torch
sys.path.insert(0, os.getcwd())
import gpd_b200
from gpd_b200.envs import HoverAviary
from gpd_b200.utils.enums import ActionType, DroneModel
E = 2097152
g = torch.Generator(device="cuda"); g.manual_seed(0)
env = HoverAviary(num_envs=E, drone_model=DroneModel.CF2P, ctrl_freq=48, act=ActionType.PID, precision="f32", auto_reset=True)
acts = [torch.rand((E, 1, 3), generator=g, device="cuda") * 2 - 1 for _ in range(4)]
env.reset()
for k in range(12):
    env._sim.step(acts[k % 4])
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 2 -f -o $O/final/prof_c5_bulk python /tmp/one_cfg.py > $O/final/ncu_c5.log 2>&1; echo "c5 rc=$?"
