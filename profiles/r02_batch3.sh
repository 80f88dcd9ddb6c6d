#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -40 > $O/pytest_b3.log
tail -15 $O/pytest_b3.log
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b3.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b3.jsonl
}
for d in 0 1 2; do
  run "direct=$d K=20" GPD_BULK_DIRECT=$d -- --steps 20 --warmup 5
  run "direct=$d K=200" GPD_BULK_DIRECT=$d -- --steps 200 --warmup 5
done
run "direct=2 tpb=128 K=20" GPD_BULK_DIRECT=2 -- --steps 20 --warmup 5 --tpb 128
run "direct=2 tpb=128 K=200" GPD_BULK_DIRECT=2 -- --steps 200 --warmup 5 --tpb 128
run "direct=0 again K=20" GPD_BULK_DIRECT=0 -- --steps 20 --warmup 5
run "direct=0 again K=200" GPD_BULK_DIRECT=0 -- --steps 200 --warmup 5
for d in 0 2; do
  GPD_BULK_DIRECT=$d timeout 200 python profiles/timeline.py 65536 0 8 > $O/timeline_b3_direct$d.txt 2>&1
done
cat $O/timeline_b3_direct0.txt
