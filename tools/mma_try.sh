export TABLE_ROWS=3
python tools/parity_table.py mnist_spring_color 100
python tools/parity_table.py mnist_spring_color 100 '{"seed": 1}'
python tools/parity_table.py spring_color 100
python tools/bench_tasks.py spring_color mnist_spring_color --profile | cut -c1-420
