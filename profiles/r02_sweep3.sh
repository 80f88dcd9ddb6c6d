#!/usr/bin/env bash
OUT=gpurun_out/r02/sweep3.jsonl
mkdir -p gpurun_out/r02
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" python bench.py --no-extra --no-cpu "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $OUT <<'PY'
import json, sys
d = json.loads(sys.argv[2])
print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                  "trials_us": [round(1e3 * t / d["steps"], 3) for t in d["trials_ms"]], "steps": d["steps"],
                  "e2e": round(d["e2e"]["value"] / 1e9, 3), "e2e_us": round(1e3 * d["e2e"]["ms_per_step"], 1)}))
PY
  tail -1 $OUT
}
run "K=20 single-graph, mirror chunks 4" -- --steps 20 --warmup 5 --launch single --e2e-steps 300
run "K=200 single, chunks 1" GPD_MIRROR_CHUNKS=1 -- --steps 200 --warmup 5 --launch single --e2e-steps 300
run "K=200 single, chunks 2" GPD_MIRROR_CHUNKS=2 -- --steps 200 --warmup 5 --launch single --e2e-steps 300
run "K=200 single, chunks 8" GPD_MIRROR_CHUNKS=8 -- --steps 200 --warmup 5 --launch single --e2e-steps 300
run "step split 2, K=200 single" GPD_STEP_SPLIT=2 -- --steps 200 --warmup 5 --launch single --e2e-steps 3
run "step split 4, K=200 single" GPD_STEP_SPLIT=4 -- --steps 200 --warmup 5 --launch single --e2e-steps 3
run "step split 2 tpb 128, K=200 single" GPD_STEP_SPLIT=2 -- --steps 200 --warmup 5 --launch single --e2e-steps 3 --tpb 128
