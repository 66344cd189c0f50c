# kappa calibration of the tcgen05 UNet forward (chains of 3 / 6 / 9 / 12 MMAs): whole-step parity margins (run under gpurun)
export TABLE_ROWS=2
for s in 0.65 0.70 0.75; do
  K=$(python -c "s=$s; print('0.27,%.3f,%.3f,%.3f' % (1.0*s, 1.75*s, 2.5*s))")
  PAIG_UNET_TC_KAPPA=$K timeout 300 python tools/parity_table.py spring_color 100
  PAIG_UNET_TC_KAPPA=$K timeout 300 python tools/parity_table.py spring_color 260 '{"seed": 1}'
  PAIG_UNET_TC_KAPPA=$K timeout 300 python tools/parity_table.py bouncing_balls 100 '{"alpha": 2.0}'
  PAIG_UNET_TC_KAPPA=$K timeout 300 python tools/parity_table.py spring_color 7 '{"alt_vel": true, "seed": 2}'
done
