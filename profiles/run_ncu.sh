#!/bin/bash
# Profiling recipe (run under gpurun from the repo root; see /opt/skills/guides/B200_PROFILING.md).
#   1) the plain command must exit 0   2) launch list with device times   3) full captures of the hot kernels
# Reports are exported to CSV on the box (gpurun_out/ is capped at 64 MiB) and the .ncu-rep kept only if small.
set -e
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
# conv3x3_kernel: 23 launches per step (12 forward c1..c12, then 11 data-gradient c12..c2); step 4 starts at 69
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 69 -c 23 -f -o gpurun_out/${TAG}_conv $CMD > gpurun_out/ncu_full1.log 2>&1
# conv3x3_wgrad_kernel: 12 per step (c12..c1)
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 36 -c 12 -f -o gpurun_out/${TAG}_wgrad $CMD > gpurun_out/ncu_full2.log 2>&1
# sgemm: 18 per step (encoder l1,l2,l3 + velocity l1,l2,l3 forward, then 12 backward); step 4 starts at 54
ncu --set full --clock-control none --import-source on -k regex:sgemm_kernel -s 54 -c 18 -f -o gpurun_out/${TAG}_sgemm $CMD > gpurun_out/ncu_full3.log 2>&1
for r in conv wgrad sgemm; do
  ncu -i gpurun_out/${TAG}_${r}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_${r}.ncu-rep --page source --csv > gpurun_out/${TAG}_${r}_source.csv 2>/dev/null || true
  sz=$(stat -c %s gpurun_out/${TAG}_${r}.ncu-rep)
  if [ "$sz" -gt 20000000 ]; then rm gpurun_out/${TAG}_${r}.ncu-rep; fi
done
gzip -f gpurun_out/${TAG}_*_source.csv || true
du -sh gpurun_out
