#!/bin/bash
# Round-1 closing captures (r1g): launch list of one spring_color step and full captures of the kernels changed since
# r1f -- fused forward / backward-data (batched gate loads, last-warp weight prefetch), TMA weight gradients -- plus
# the other tasks' new paths: 3bp's fused backward at 36 px and its cp.async per-layer weight gradients, mnist's
# channel-blocked TMA weight gradients and its per-layer conv.  Run under gpurun from the repo root.
set -e
TAG=r1g
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 100 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
cap() {  # name regex skip count [command]
  local C="${5:-$CMD}"
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/${TAG}_$1 $C > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_$1_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap fused_fwd unet_fused_fwd 3 1
cap fused_bwd unet_fused_bwd 3 1
cap wgrad wgrad_tma 36 12
cap 3bp_fused_bwd unet_fused_bwd 3 1 "python tools/bench_tasks.py 3bp_color"
cap 3bp_wgrad "conv3x3_wgrad" 36 12 "python tools/bench_tasks.py 3bp_color"
cap 3bp_decode "decode_kernel" 3 1 "python tools/bench_tasks.py 3bp_color"
cap mnist_wgrad wgrad_tma 51 17 "python tools/bench_tasks.py mnist_spring_color"
cap mnist_conv "conv3x3_kernel" 102 34 "python tools/bench_tasks.py mnist_spring_color"
du -sh gpurun_out
