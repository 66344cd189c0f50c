#!/bin/bash
# r1i: the fused UNet kernels after the instruction-footprint change (compare with r1g: issue %, no_instruction stall)
set -e
TAG=r1i
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
cap() {
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/${TAG}_$1 $CMD > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_$1_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap fused_fwd unet_fused_fwd 3 1
cap fused_bwd unet_fused_bwd 3 1
