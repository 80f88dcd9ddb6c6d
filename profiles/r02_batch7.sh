#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b7.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b7.jsonl
}
run "prefetch under the claim K=200" -- --steps 200 --warmup 5
run "no prefetch K=200" GPD_DEBUG_UNSAFE=4 -- --steps 200 --warmup 5
run "UNSAFE relaxed publish K=200" GPD_DEBUG_UNSAFE=8 -- --steps 200 --warmup 5
run "UNSAFE relaxed publish, no prefetch K=200" GPD_DEBUG_UNSAFE=12 -- --steps 200 --warmup 5
run "UNSAFE no claim/publish K=200" GPD_DEBUG_UNSAFE=3 -- --steps 200 --warmup 5
run "prefetch under the claim K=20" -- --steps 20 --warmup 5
run "no prefetch K=20" GPD_DEBUG_UNSAFE=4 -- --steps 20 --warmup 5
run "prefetch direct=0 K=200" GPD_BULK_DIRECT=0 -- --steps 200 --warmup 5
run "prefetch f64 K=200" -- --steps 200 --warmup 5 --precision f64
run "no prefetch f64 K=200" GPD_DEBUG_UNSAFE=4 -- --steps 200 --warmup 5 --precision f64
timeout 200 python profiles/timeline.py 65536 0 8 > $O/timeline_b7.txt 2>&1; cat $O/timeline_b7.txt
