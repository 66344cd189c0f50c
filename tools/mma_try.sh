python -m pytest tests/test_gpu_stages.py -x -q -k "conv3x3_primitive" 2>&1 | tail -2
PAIG_PROFILE_LAYERS=1 python tools/bench_tasks.py spring_color --profile | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); k=d['kernels_ms_per_step']
print(d['ms_per_step'], 'MMA', {a:b for a,b in k.items() if 'wgrad' in a})"
python tools/bench_tasks.py spring_color 3bp_color mnist_spring_color | cut -c1-330
