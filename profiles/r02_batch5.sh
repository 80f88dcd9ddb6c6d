#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_b5.log
tail -5 $O/pytest_b5.log
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b5.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b5.jsonl
}
run "default K=20" -- --steps 20 --warmup 5
run "default K=200" -- --steps 200 --warmup 5
run "direct=0 K=20" GPD_BULK_DIRECT=0 -- --steps 20 --warmup 5
run "tpb=128 K=20" -- --steps 20 --warmup 5 --tpb 128
run "tpb=96 K=20" -- --steps 20 --warmup 5 --tpb 96
run "tpb=32 K=20" -- --steps 20 --warmup 5 --tpb 32
run "sets=4 K=20" -- --steps 20 --warmup 5 --sets 4
run "48Hz K=200" -- --steps 200 --warmup 5 --ctrl-freq 48 --sets 6
for c in 1 2 4 8; do
  GPD_MIRROR_CHUNKS=$c python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --e2e-steps 300 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
done
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_b5_full.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_b5_full.json"))
print("headline", d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"])
for k, v in d["other_configs"].items():
    print(k, v.get("us_per_step"), v.get("roofline", {}).get("frac"), v.get("error"))
PY
timeout 200 python profiles/timeline.py 65536 0 8 > $O/timeline_b5.txt 2>&1; cat $O/timeline_b5.txt
