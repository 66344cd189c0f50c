#!/bin/bash
# r2s (closing set of round 2): launch list of one spring_color step, --set full of the tcgen05 ShallowUNet forward
# (csrc/unet_tc.cu), the fused backward-data kernel and ALL 12 weight-gradient launches of one step.
# Every command runs once without ncu first (B200_PROFILING.md); numbers printed under ncu are never bench values.
TAG=r2s
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_${name}_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 135 -c 50 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_${TAG}_launches.log 2>&1
cap tc_fwd unet_tc_fwd 3 1 $B
cap fused_bwd unet_fused_bwd 3 1 $B
cap wgrad "conv3x3_wgrad" 36 12 $B
