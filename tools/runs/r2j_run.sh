python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -8
python tools/tune_slos.py 12 24 2>&1 | tail -1
python tools/time_slab.py 12 24 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('bench ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'e2e', d['e2e']['ms_per_step'])"
