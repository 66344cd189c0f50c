T="mnist_spring_color 100"
export TABLE_ROWS=4
python tools/parity_table.py $T
PAIG_CONV_TC_KAPPA=0.36 python tools/parity_table.py $T
PAIG_CONV_TC_KAPPA=0.18 python tools/parity_table.py $T
PAIG_CONV_TC_DRAIN=3 python tools/parity_table.py $T
python tools/parity_table.py mnist_spring_color 16
python tools/bench_tasks.py mnist_spring_color --profile
PAIG_CONV_TC_DRAIN=3 python tools/bench_tasks.py mnist_spring_color
python tools/conv_tc_probe.py big
