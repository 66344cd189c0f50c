# kappa sweep of the tcgen05 UNet forward + whole-step parity tables (run under gpurun)
export COMPACT=1 TABLE_ROWS=5
for k in 0 0.3 0.62; do PAIG_UNET_TC_KAPPA=$k timeout 200 python tools/unet_tc_check.py 8; done
for k in 0.62 0.3 0; do PAIG_UNET_TC=1 PAIG_UNET_TC_KAPPA=$k timeout 300 python tools/parity_table.py spring_color 100; done
