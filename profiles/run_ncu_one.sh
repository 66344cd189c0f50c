#!/bin/bash
# One full ncu capture of one kernel of the bench step:  bash profiles/run_ncu_one.sh <tag> <kernel-regex> <skip> <count>
set -e
TAG=$1; KRE=$2; SKIP=${3:-3}; CNT=${4:-1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -f -o gpurun_out/${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null || true
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null || true
gzip -f gpurun_out/${TAG}_source.csv || true
sz=$(stat -c %s gpurun_out/${TAG}.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/${TAG}.ncu-rep; fi
du -sh gpurun_out
