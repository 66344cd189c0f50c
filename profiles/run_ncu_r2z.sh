#!/bin/bash
# r2z (closing set of round 2): launch list of one spring_color step; --set full of the tcgen05 ShallowUNet kernel under its
# forward and its backward-data plan (both are launches of unet_tc_fwd_kernel), of ALL 12 weight-gradient launches of one spring
# step, and of the tcgen05 weight-gradient kernel (csrc/wgrad_tc.cu) on three layers of the 64-px UNet at B = 100.
# Every command runs once without ncu first (B200_PROFILING.md); numbers printed under ncu are never bench values.
TAG=r2z
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
W="python tools/wgrad_tc_probe.py --run [[1000,32,32,64],[1000,64,32,32],[1000,128,128,8]]"
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
$W > gpurun_out/${TAG}_plain_wgrad_tc.log 2>&1 &&
# per shape: 1 correctness launch + 3 warm-up + 10 timed = 14 launches; take the 5th launch of each shape
for k in 0 1 2; do cap wgrad_tc_$k conv3x3_wgrad_tc $((14 * k + 4)) 1 $W; done
