#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_b15.log
tail -5 $O/pytest_b15.log
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b15.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b15.jsonl
}
run "default K=20" -- --steps 20 --warmup 5
run "default K=200" -- --steps 200 --warmup 5
run "default 1M K=48" -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "default 32768 envs K=200" -- --steps 200 --warmup 5 --envs 32768 --sets 16
run "default 131072 envs K=100" -- --steps 100 --warmup 5 --envs 131072 --sets 6
LABEL="default" python profiles/r02_others.py 2>/dev/null | tail -1
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_b15_full.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_b15_full.json"))
print("headline", d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"])
print("async", d["async_pools"]["ms_per_step"], "l2_resident", d["l2_resident"]["ms_per_step"])
for k, v in d["other_configs"].items():
    print(k, v.get("us_per_step"), v.get("roofline", {}).get("frac"), v.get("error"))
PY
timeout 200 python profiles/timeline.py 65536 0 8 > $O/timeline_b15.txt 2>&1; cat $O/timeline_b15.txt
