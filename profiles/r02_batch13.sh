#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
for tpc in 2 1; do
  GPD_BULK_TPC=$tpc timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "sequencing or chained or bulk or full_size or bench_contract or fuzz" 2>&1 | tail -3
done
run() {
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  line=$(env "${envs[@]}" timeout 300 python bench.py --no-extra --no-cpu --e2e-steps 3 --trials 15 "$@" 2>/dev/null | tail -1)
  python - "$label" "$line" >> $O/sweep_b13.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b13.jsonl
}
for tpc in 1 2 3 4; do
  run "tpc=$tpc K=200" GPD_BULK_TPC=$tpc -- --steps 200 --warmup 5
  run "tpc=$tpc K=20" GPD_BULK_TPC=$tpc -- --steps 20 --warmup 5
done
run "tpc=2 tpb=32 K=200" GPD_BULK_TPC=2 -- --steps 200 --warmup 5 --tpb 32
run "tpc=4 tpb=32 K=200" GPD_BULK_TPC=4 -- --steps 200 --warmup 5 --tpb 32
run "tpc=2 tpb=128 K=200" GPD_BULK_TPC=2 -- --steps 200 --warmup 5 --tpb 128
run "tpc=2 f64 K=200" GPD_BULK_TPC=2 -- --steps 200 --warmup 5 --precision f64
run "tpc=1 f64 K=200" GPD_BULK_TPC=1 -- --steps 200 --warmup 5 --precision f64
for tpc in 1 2; do
GPD_BULK_TPC=$tpc python bench.py --steps 20 --warmup 5 --no-cpu --no-others 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tpc $tpc: headline', d['ms_per_step'], 'async', d['async_pools']['ms_per_step'], 'l2_resident (one set, dependent steps)', d['l2_resident']['ms_per_step'])"
done
