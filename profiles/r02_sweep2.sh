#!/usr/bin/env bash
# Round-2 sweep 2 (run under gpurun): launch modes of the bench window and the no-DMA-warp kernel variant.
OUT=gpurun_out/r02/sweep2.jsonl
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
                  "trials_us": [round(1e3 * t / d["steps"], 3) for t in d["trials_ms"]], "steps": d["steps"], "launch": d["config"]["launch"][:40]}))
PY
  tail -1 $OUT
}
run "dma-warp K=20 direct" -- --steps 20 --warmup 5
run "dma-warp K=20 single-graph" -- --steps 20 --warmup 5 --launch single
run "dma-warp K=200 direct" -- --steps 200 --warmup 5
run "dma-warp K=1000 single" -- --steps 1000 --warmup 5
run "dma-warp K=20000 cycles(128)" -- --steps 20000 --warmup 5
for tpb in 64 96 128; do
  run "inline-tma tpb=$tpb K=20 direct" GPD_DMA_WARP=0 -- --steps 20 --warmup 5 --tpb $tpb
  run "inline-tma tpb=$tpb K=200 direct" GPD_DMA_WARP=0 -- --steps 200 --warmup 5 --tpb $tpb
done
run "inline-tma tpb=64 no-edge K=200" GPD_DMA_WARP=0 GPD_TMA_EDGE=0 -- --steps 200 --warmup 5 --tpb 64
run "inline-tma tpb=64 48Hz K=200" GPD_DMA_WARP=0 -- --steps 200 --warmup 5 --ctrl-freq 48 --sets 6
run "inline-tma tpb=64 f64 K=200" GPD_DMA_WARP=0 -- --steps 200 --warmup 5 --precision f64
run "inline-tma tpb=128 f64 K=200" GPD_DMA_WARP=0 -- --steps 200 --warmup 5 --precision f64 --tpb 128
run "default 1M envs K=48 (auto: serial)" -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "inline-tma 1M envs K=48" GPD_DMA_WARP=0 -- --steps 48 --warmup 5 --envs 1048576 --sets 2
run "inline-tma 262144 envs K=96" GPD_DMA_WARP=0 -- --steps 96 --warmup 5 --envs 262144 --sets 4
