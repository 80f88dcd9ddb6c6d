#!/usr/bin/env bash
OUT=gpurun_out/r02/sweep4.jsonl
mkdir -p gpurun_out/r02
run() {
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
run "bulk K=20 single" -- --steps 20 --warmup 5 --launch single
run "bulk K=20 direct" -- --steps 20 --warmup 5 --launch direct
run "bulk K=200 single" -- --steps 200 --warmup 5 --launch single
run "bulk K=20000" -- --steps 20000 --warmup 5
run "legacy K=200 single" GPD_BULK=0 -- --steps 200 --warmup 5 --launch single
run "bulk tpb=32 K=200" -- --steps 200 --warmup 5 --launch single --tpb 32
run "bulk tpb=96 K=200" -- --steps 200 --warmup 5 --launch single --tpb 96
run "bulk tpb=128 K=200" -- --steps 200 --warmup 5 --launch single --tpb 128
run "bulk serial (no tile_dep/PDL) K=200" GPD_TILE_DEP=0 GPD_PDL=0 -- --steps 200 --warmup 5 --launch single
run "bulk 48Hz K=200" -- --steps 200 --warmup 5 --launch single --ctrl-freq 48 --sets 6
run "bulk f64 K=200" -- --steps 200 --warmup 5 --launch single --precision f64
run "bulk f64 tpb=64 K=200" -- --steps 200 --warmup 5 --launch single --precision f64 --tpb 64
run "bulk 262144 envs K=96" -- --steps 96 --warmup 5 --launch single --envs 262144 --sets 4
run "bulk 1M envs K=48" -- --steps 48 --warmup 5 --launch single --envs 1048576 --sets 2
run "bulk 1M envs K=48 tile_dep" GPD_TILE_DEP=1 -- --steps 48 --warmup 5 --launch single --envs 1048576 --sets 2
run "legacy 1M envs K=48" GPD_BULK=0 -- --steps 48 --warmup 5 --launch single --envs 1048576 --sets 2
