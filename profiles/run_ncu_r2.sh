#!/bin/bash
# r2: launch list of one spring_color step (45 -> 44 launches, side streams serialised by ncu), --set full of the three
# ShallowUNet kernels, and of the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) inside a mnist step.
# Every command runs once without ncu first (B200_PROFILING.md); numbers printed under ncu are never bench values.
TAG=r2
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
M="python tools/bench_tasks.py mnist_spring_color --steps 1"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_${name}_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 132 -c 50 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_${TAG}_launches.log 2>&1
cap fused_fwd unet_fused_fwd 3 1 $B
cap fused_bwd unet_fused_bwd 3 1 $B
cap wgrad_tma conv3x3_wgrad_tma 36 3 $B
$M > gpurun_out/${TAG}_plain_mnist.log 2>&1 &&
cap conv_tc conv3x3_tc_kernel 60 6 $M
