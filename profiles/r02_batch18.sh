#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/pytest_b18.log
tail -8 $O/pytest_b18.log
LABEL="multi bulk" python profiles/r02_others.py 2>/dev/null | tail -1
LABEL="no bulk" GPD_BULK=0 python profiles/r02_others.py 2>/dev/null | tail -1
timeout 300 python profiles/configs.py c3_multihover2_gnd_drag_f64 c3_multihover2_gnd_drag_f32 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], round(d['us_per_step'], 2), round(d.get('frac_of_hbm_peak', 0), 3))"
GPD_BULK=0 timeout 300 python profiles/configs.py c3_multihover2_gnd_drag_f32 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('no bulk', d['config'], round(d['us_per_step'], 2), round(d.get('frac_of_hbm_peak', 0), 3))"
