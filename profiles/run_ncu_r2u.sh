#!/bin/bash
# r2u (closing set of round 2): launch list of one spring_color step, --set full of the tcgen05 ShallowUNet kernel under its
# forward and its backward-data plan (csrc/unet_tc.cu; both are launches of unet_tc_fwd_kernel) and ALL 12 weight-gradient
# launches of one step.  Every command runs once without ncu first (B200_PROFILING.md); numbers printed under ncu are never
# bench values.
TAG=r2u
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
# launches of unet_tc_fwd_kernel alternate forward, backward-data: skip 3 steps = 6 launches, take the next two
cap tc_fwd_bwd unet_tc_fwd 6 2 $B
cap wgrad "conv3x3_wgrad" 36 12 $B
