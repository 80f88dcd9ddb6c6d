#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/pytest_b11.log
tail -6 $O/pytest_b11.log
for zc in 1 0; do
  GPD_MIRROR_ZEROCOPY=$zc python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --e2e-steps 400 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('zero_copy $zc e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_obs_equals_device_obs'])"
done
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_b11_full.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_b11_full.json"))
print("headline", d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d.get("e2e_pools"))
PY
