#!/usr/bin/env bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one --set full capture of the
# step kernel. Run under gpurun on ONE GPU; reports land in gpurun_out/ and summaries are copied to profiles/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 64 --warmup 16 --no-cpu --no-graph --e2e-steps 2 ${BENCH_EXTRA:-}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 32 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 40 -c 3 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
