#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_b9.log
tail -6 $O/pytest_b9.log
LABEL="bulk A<4" python profiles/r02_others.py 2>/dev/null | tail -1
LABEL="no bulk" GPD_BULK=0 python profiles/r02_others.py 2>/dev/null | tail -1
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > $O/bench_b9_full.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_b9_full.json"))
print("headline", d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d.get("e2e_pools"))
for k, v in d["other_configs"].items():
    print(k, v.get("us_per_step"), v.get("roofline", {}).get("frac"), v.get("error"))
PY
