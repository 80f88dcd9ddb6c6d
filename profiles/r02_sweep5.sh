#!/usr/bin/env bash
OUT=gpurun_out/r02/sweep5.jsonl
mkdir -p gpurun_out/r02
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "trials_us": [round(1e3 * t / d["steps"], 3) for t in d["trials_ms"]], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $OUT
}
run "bulk v1 K=200" -- --steps 200 --warmup 5
run "pipe 2 tiles K=200" GPD_PIPE=1 -- --steps 200 --warmup 5
run "pipe 2 tiles K=20" GPD_PIPE=1 -- --steps 20 --warmup 5
run "pipe 3 tiles K=200" GPD_PIPE=1 GPD_PIPE_TILES=3 -- --steps 200 --warmup 5
run "pipe 4 tiles K=200" GPD_PIPE=1 GPD_PIPE_TILES=4 -- --steps 200 --warmup 5
run "pipe 1 tile K=200" GPD_PIPE=1 GPD_PIPE_TILES=1 -- --steps 200 --warmup 5
run "pipe serial (no tile_dep) K=200" GPD_PIPE=1 GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 200 --warmup 5
run "pipe 1M envs K=48" GPD_PIPE=1 -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "pipe 262144 envs K=96" GPD_PIPE=1 -- --steps 96 --warmup 5 --envs 262144 --sets 4
