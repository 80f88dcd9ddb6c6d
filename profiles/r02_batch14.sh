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
  python - "$label" "$line" >> $O/sweep_b14.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    t = sorted(round(1e3 * t / d["steps"], 3) for t in d["trials_ms"])
    print(json.dumps({"label": sys.argv[1], "us_per_step": round(1e3 * d["ms_per_step"], 3), "frac": round(d["roofline"]["frac"], 4),
                      "min": t[0], "max": t[-1], "steps": d["steps"]}))
except Exception as ex:
    print(json.dumps({"label": sys.argv[1], "error": repr(ex)[:100]}))
PY
  tail -1 $O/sweep_b14.jsonl
}
for tpb in 96 128; do
  for tpc in 1 2 3; do
    run "tpb=$tpb tpc=$tpc K=200" GPD_BULK_TPC=$tpc -- --steps 200 --warmup 5 --tpb $tpb
    run "tpb=$tpb tpc=$tpc K=20" GPD_BULK_TPC=$tpc -- --steps 20 --warmup 5 --tpb $tpb
  done
done
run "tpb=128 tpc=2 48Hz K=200" GPD_BULK_TPC=2 -- --steps 200 --warmup 5 --tpb 128 --ctrl-freq 48 --sets 6
run "tpb=64 tpc=1 48Hz K=200" GPD_BULK_TPC=1 -- --steps 200 --warmup 5 --ctrl-freq 48 --sets 6
run "tpb=128 tpc=2 f64 K=200" GPD_BULK_TPC=2 -- --steps 200 --warmup 5 --tpb 128 --precision f64
run "tpb=128 tpc=2 262144 envs K=96" GPD_BULK_TPC=2 -- --steps 96 --warmup 5 --tpb 128 --envs 262144 --sets 4
run "tpb=64 tpc=1 262144 envs K=96" GPD_BULK_TPC=1 -- --steps 96 --warmup 5 --envs 262144 --sets 4
run "tpb=128 tpc=2 16384 envs K=200" GPD_BULK_TPC=2 -- --steps 200 --warmup 5 --tpb 128 --envs 16384 --sets 32
run "tpb=64 tpc=1 16384 envs K=200" GPD_BULK_TPC=1 -- --steps 200 --warmup 5 --envs 16384 --sets 32
for cfg in "1 0" "2 128"; do
set -- $cfg
GPD_BULK_TPC=$1 python bench.py --steps 20 --warmup 5 --no-cpu --no-others --tpb $2 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tpc $1 tpb $2: headline', d['ms_per_step'], 'async', d['async_pools']['ms_per_step'], 'l2_resident (one set, dependent steps)', d['l2_resident']['ms_per_step'])"
done
