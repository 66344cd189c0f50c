# A/B of the tcgen05 backward-data pass (PAIG_UNET_TC_BWD=1) against the FMA kernel: whole-step parity tables + timing
export TABLE_ROWS=3
PAIG_UNET_TC_BWD=1 timeout 300 python tools/parity_table.py spring_color 3
PAIG_UNET_TC_BWD=1 timeout 300 python tools/parity_table.py spring_color 100
PAIG_UNET_TC_BWD=1 timeout 300 python tools/bench_tasks.py spring_color --profile 2>&1 | cut -c1-900
