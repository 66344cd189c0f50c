#!/bin/bash
# r2n: --set full of the tensor-core (mma.sync 3xTF32) variants of the fused ShallowUNet kernels inside a spring_color step.
TAG=${TAG:-r2n}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_${name}_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 || exit 1
cap fused_fwd unet_fused_fwd 3 1 $B
[ -n "$FWD_ONLY" ] || cap fused_bwd unet_fused_bwd 3 1 $B
