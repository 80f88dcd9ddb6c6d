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
  python - "$label" "$line" >> $O/sweep_b8.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b8.jsonl
}
run "no prefetch (4) K=200" GPD_DEBUG_UNSAFE=4 -- --steps 200 --warmup 5
run "stats after publish (4+16) K=200" GPD_DEBUG_UNSAFE=20 -- --steps 200 --warmup 5
run "stats after publish direct=0 (4+16) K=200" GPD_DEBUG_UNSAFE=20 GPD_BULK_DIRECT=0 -- --steps 200 --warmup 5
run "direct=0 (4) K=200" GPD_DEBUG_UNSAFE=4 GPD_BULK_DIRECT=0 -- --steps 200 --warmup 5
run "UNSAFE relaxed publish direct=0 (4+8) K=200" GPD_DEBUG_UNSAFE=12 GPD_BULK_DIRECT=0 -- --steps 200 --warmup 5
run "stats after publish (4+16) K=20" GPD_DEBUG_UNSAFE=20 -- --steps 20 --warmup 5
run "no prefetch (4) K=20" GPD_DEBUG_UNSAFE=4 -- --steps 20 --warmup 5
run "tpb=128 (4+16) K=200" GPD_DEBUG_UNSAFE=20 -- --steps 200 --warmup 5 --tpb 128
run "tpb=128 (4) K=200" GPD_DEBUG_UNSAFE=4 -- --steps 200 --warmup 5 --tpb 128
