# kappa calibration of the tcgen05 UNet forward (chains of 3 / 6 / 9 / 12 MMAs) + whole-step parity tables (run under gpurun)
export COMPACT=1 TABLE_ROWS=4
for s in 0.5 0.6 0.7 0.8; do
  K=$(python -c "s=$s; print('0.27,%.3f,%.3f,%.3f' % (1.0*s, 1.75*s, 2.5*s))")
  PAIG_UNET_TC_KAPPA=$K timeout 200 python tools/unet_tc_check.py 30 | grep -v "^FMA\|^TC"
done
for s in 0.6 0.7 0.8; do
  K=$(python -c "s=$s; print('0.27,%.3f,%.3f,%.3f' % (1.0*s, 1.75*s, 2.5*s))")
  PAIG_UNET_TC_KAPPA=$K timeout 300 python tools/parity_table.py spring_color 100
done
