#!/usr/bin/env bash
# A/B of the bulk kernel's shared-memory footprint (GPD_BULK_DIRECT) and the one-word tile sequencing; 15 windows per point
set -u
O=gpurun_out/r02
mkdir -p $O
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b2.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b2.jsonl
}
for d in 0 1 2; do
  GPD_BULK_DIRECT=$d timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
done
for d in 0 1 2; do
  run "direct=$d K=20" GPD_BULK_DIRECT=$d -- --steps 20 --warmup 5
  run "direct=$d K=200" GPD_BULK_DIRECT=$d -- --steps 200 --warmup 5
done
for tpb in 32 96 128; do
  run "direct=1 tpb=$tpb K=20" GPD_BULK_DIRECT=1 -- --steps 20 --warmup 5 --tpb $tpb
  run "direct=2 tpb=$tpb K=20" GPD_BULK_DIRECT=2 -- --steps 20 --warmup 5 --tpb $tpb
  run "direct=2 tpb=$tpb K=200" GPD_BULK_DIRECT=2 -- --steps 200 --warmup 5 --tpb $tpb
done
run "direct=0 1M K=48" GPD_BULK_DIRECT=0 -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "direct=1 1M K=48" GPD_BULK_DIRECT=1 -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "direct=2 1M K=48" GPD_BULK_DIRECT=2 -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "direct=0 f64 K=200" GPD_BULK_DIRECT=0 -- --steps 200 --warmup 5 --precision f64
run "direct=1 f64 K=200" GPD_BULK_DIRECT=1 -- --steps 200 --warmup 5 --precision f64
run "direct=2 f64 K=200" GPD_BULK_DIRECT=2 -- --steps 200 --warmup 5 --precision f64
for d in 0 1 2; do
  GPD_BULK_DIRECT=$d timeout 200 python profiles/timeline.py 65536 0 8 > $O/timeline_b2_direct$d.txt 2>&1
done
python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > $O/bench_b2_full.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_b2_full.json"))
for k, v in d["other_configs"].items():
    print(k, v.get("us_per_step"), v.get("roofline", {}).get("frac"))
PY
