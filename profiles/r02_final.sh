#!/usr/bin/env bash
# Final round-2 evidence with the shipped build: plain bench, ncu launch list of the same command (graph nodes), ncu --set full of
# the dominant kernel at the bench size and at 1 M envs, and of the C5 shape on the bulk kernel; per-CTA timeline.
set -u
O=gpurun_out/r02/final
mkdir -p $O
CMD="python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --e2e-steps 2"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 600 --csv --log-file $O/launches_bench_k20_graph.csv $CMD > $O/ncu_launch.log 2>&1
echo "launch-list rc=$?"
CMD2="python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --no-graph --e2e-steps 2"
$CMD2 > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -f -o $O/prof_bulk_65536 $CMD2 > $O/ncu_full.log 2>&1
echo "full rc=$?"
CMD3="python bench.py --steps 6 --warmup 3 --no-cpu --no-extra --no-graph --e2e-steps 2 --envs 1048576 --sets 2"
ncu --set full --clock-control none -k regex:step_kernel -s 10 -c 2 -f -o $O/prof_bulk_1M $CMD3 > $O/ncu_full_1M.log 2>&1
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
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 2 -f -o $O/prof_c5_bulk python /tmp/one_cfg.py > $O/ncu_c5.log 2>&1; echo "c5 rc=$?"
timeout 200 python profiles/timeline.py 65536 0 8 > $O/timeline_final.txt 2>&1; cat $O/timeline_final.txt
python bench.py --steps 200 --warmup 5 --no-cpu --no-extra --e2e-steps 2 --trials 15 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('K=200', d['ms_per_step'], d['roofline']['frac'])"
ls -la $O
