#!/bin/bash
# r2o: --set full of the tensor-core weight-gradient kernel (mma.sync 3xTF32 on the TMA strip pipeline), all 12 launches of a spring step.
TAG=${TAG:-r2o}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_${name}_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$name.ncu-rep
}
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 || exit 1
cap wgrad_mma conv3x3_wgrad_mma 36 12 $B
