#!/usr/bin/env bash
# round-2 evidence batch: timeline of the bulk kernel, block-size sweep at the driver's K=20, ncu launch list (graph nodes) and one
# --set full capture of the dominant kernel.  Run under gpurun on ONE GPU; everything lands in gpurun_out/r02/.
set -u
O=gpurun_out/r02
mkdir -p $O
for tpb in 0 32 128; do
  timeout 200 python profiles/timeline.py 65536 $tpb 8 > $O/timeline_bulk_tpb$tpb.txt 2>&1
done
GPD_BULK=0 timeout 200 python profiles/timeline.py 65536 0 8 > $O/timeline_tmabox.txt 2>&1
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b1.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "trials_us": [round(1e3 * t / d["steps"], 3) for t in d["trials_ms"]], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b1.jsonl
}
run "default K=20" -- --steps 20 --warmup 5
run "default K=200" -- --steps 200 --warmup 5
run "tpb=32 K=20" -- --steps 20 --warmup 5 --tpb 32
run "tpb=96 K=20" -- --steps 20 --warmup 5 --tpb 96
run "tpb=128 K=20" -- --steps 20 --warmup 5 --tpb 128
run "tpb=128 K=200" -- --steps 200 --warmup 5 --tpb 128
run "no tile_dep/PDL K=200" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 200 --warmup 5
run "tma-box kernel K=200" GPD_BULK=0 -- --steps 200 --warmup 5
run "1M envs K=48" -- --steps 48 --warmup 5 --envs 1048576 --sets 2
CMD="python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --e2e-steps 2"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 400 --csv --log-file $O/launches_bench_k20_graph.csv $CMD > $O/ncu_launch.log 2>&1
echo "launch-list rc=$?"
CMD2="python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --no-graph --e2e-steps 2"
$CMD2 > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -f -o $O/prof_bulk_65536 $CMD2 > $O/ncu_full.log 2>&1
echo "full rc=$?"
CMD3="python bench.py --steps 6 --warmup 3 --no-cpu --no-extra --no-graph --e2e-steps 2 --envs 1048576 --sets 2"
ncu --set full --clock-control none -k regex:step_kernel -s 10 -c 2 -f -o $O/prof_bulk_1M $CMD3 > $O/ncu_full_1M.log 2>&1
echo "full 1M rc=$?"
ls -la $O
