#!/bin/bash
# Round-1 profiling recipe (run under gpurun from the repo root; B200_PROFILING.md).
# 1) plain run must exit 0, 2) launch list with device times, 3) full capture of the conv kernels.
set -e
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 70 -c 35 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
