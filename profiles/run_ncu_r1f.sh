#!/bin/bash
# Round-1 final captures: launch list of one step, then full captures of the three dominant kernels, of the
# decoder / rollout kernels the north star names, and of the GEMM / MLP kernels (fp32 l1 forward, tcgen05 l1 backward,
# fused MLP tails).  Run under gpurun from the repo root.
set -e
TAG=r1f
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 100 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/${TAG}_$1 $CMD > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null || true
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_$1_source.csv.gz || true
  rm -f gpurun_out/${TAG}_$1.ncu-rep
}
cap fused_fwd unet_fused_fwd 3 1
cap fused_bwd unet_fused_bwd 3 1
cap wgrad wgrad_tma 36 12
cap decode "decode_" 6 2
cap rollout "rollout_" 6 2
cap sgemm sgemm_kernel 3 1
cap tc gemm_tf32x3 6 2
cap tail "enc_tail|vel_mlp" 12 4
du -sh gpurun_out
