B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 9 -c 3 -f -o gpurun_out/r2z_gemm $B > gpurun_out/ncu_r2z_gemm.log 2>&1
ncu -i gpurun_out/r2z_gemm.ncu-rep --page raw --csv > gpurun_out/r2z_gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/r2z_gemm.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2z_gemm_source.csv.gz
rm -f gpurun_out/r2z_gemm.ncu-rep
