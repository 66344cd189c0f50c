#!/bin/bash
# r1h: the kernels added after the r1g set -- cp.async weight gradients of 3bp's 18/9-px levels (row-per-thread
# staging, 8 channels per thread), the double-buffered cp.async conv3x3 of the mnist per-layer path, the vectorised
# upsample / gate kernels.  Run under gpurun from the repo root.
set -e
TAG=r1h
cap() {  # name regex skip count command
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/${TAG}_$1 $5 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_$1_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap 3bp_wgrad_cpasync "conv3x3_wgrad_kernel" 21 7 "python tools/bench_tasks.py 3bp_color"
cap mnist_conv_async "conv3x3_async_kernel" 60 10 "python tools/bench_tasks.py mnist_spring_color"
cap mnist_elementwise "upsample2_quad|relu_gate_vec" 30 6 "python tools/bench_tasks.py mnist_spring_color"
du -sh gpurun_out
