#!/usr/bin/env bash
# Round-2 A/B sweep of the step-kernel scheduling (run under gpurun): per-CTA sequencing + PDL (default) vs whole-grid
# stream order, block sizes, grid sizes.  Appends JSON lines to gpurun_out/r02/sweep.jsonl.
OUT=gpurun_out/r02/sweep.jsonl
mkdir -p gpurun_out/r02
run() {  # label, env assignments..., -- bench args
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" python bench.py --no-extra --no-cpu --e2e-steps 3 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $OUT <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                  "trials_us": [round(1e3 * t / d["steps"], 3) for t in d["trials_ms"]], "steps": d["steps"]}))
PY
  tail -1 $OUT
}
run "default K=20" -- --steps 20 --warmup 5
run "default K=200" -- --steps 200 --warmup 5
run "default K=20000" -- --steps 20000 --warmup 5
run "serial (TILE_DEP=0 PDL=0) K=20" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 20 --warmup 5
run "serial (TILE_DEP=0 PDL=0) K=20000" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 20000 --warmup 5
run "tile_dep without PDL K=200" GPD_PDL=0 -- --steps 200 --warmup 5
run "late trigger K=200" GPD_PDL_EARLY=0 -- --steps 200 --warmup 5
for tpb in 32 64 96 128; do run "default tpb=$tpb K=200" -- --steps 200 --warmup 5 --tpb $tpb; done
run "no TMA edge K=200" GPD_TMA_EDGE=0 -- --steps 200 --warmup 5
run "default 1M envs K=48" -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "serial 1M envs K=48" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "default 262144 envs K=96" -- --steps 96 --warmup 5 --envs 262144 --sets 4
run "serial 262144 envs K=96" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 96 --warmup 5 --envs 262144 --sets 4
run "default 48Hz K=200" -- --steps 200 --warmup 5 --ctrl-freq 48 --sets 6
run "serial 48Hz K=200" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 200 --warmup 5 --ctrl-freq 48 --sets 6
run "default f64 K=200" -- --steps 200 --warmup 5 --precision f64
run "serial f64 K=200" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 200 --warmup 5 --precision f64
