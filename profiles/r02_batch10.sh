#!/usr/bin/env bash
set -u
O=gpurun_out/r02
mkdir -p $O
LABEL="96 registers (default build)" python profiles/r02_f64_single.py 2>/dev/null | tail -1 | tee -a $O/f64_single.jsonl
GPD_NVCC_EXTRA="-DGPD_F64_SINGLE_MINB=3" bash gym-pybullet-drones-routing_b200/csrc/build.sh > /dev/null 2>&1
LABEL="128 registers (GPD_F64_SINGLE_MINB=3)" python profiles/r02_f64_single.py 2>/dev/null | tail -1 | tee -a $O/f64_single.jsonl
python profiles/ptxas_summary.py | grep "step_kernel<double, ., false" | tee $O/ptxas_f64_minb3.txt
