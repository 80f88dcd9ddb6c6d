#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_b4.log
tail -8 $O/pytest_b4.log
timeout 600 python profiles/r02_lane_split.py > $O/lane_split.jsonl 2> $O/lane_split.err; cat $O/lane_split.jsonl; tail -3 $O/lane_split.err
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b4.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b4.jsonl
}
for d in 0 2 1; do
  run "direct=$d K=20" GPD_BULK_DIRECT=$d -- --steps 20 --warmup 5
  run "direct=$d K=200" GPD_BULK_DIRECT=$d -- --steps 200 --warmup 5
done
run "direct=2 48Hz bulk K=200" GPD_BULK_DIRECT=2 GPD_BULK=1 -- --steps 200 --warmup 5 --ctrl-freq 48 --sets 6
run "48Hz default K=200" -- --steps 200 --warmup 5 --ctrl-freq 48 --sets 6
run "direct=2 f64 K=200" GPD_BULK_DIRECT=2 -- --steps 200 --warmup 5 --precision f64
run "direct=0 f64 K=200" GPD_BULK_DIRECT=0 -- --steps 200 --warmup 5 --precision f64
GPD_BULK_DIRECT=2 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > $O/bench_b4_full.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_b4_full.json"))
print("headline", d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k, v in d["other_configs"].items():
    print(k, v.get("us_per_step"), v.get("roofline", {}).get("frac"), v.get("error"))
PY
