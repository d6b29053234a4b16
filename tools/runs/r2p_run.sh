python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/tune_slos.py 12 24 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -c 300 gpurun_out/r2p_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().split('\n')[-1])
print('ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
PY
python __graft_entry__.py smoke 2>&1 | tail -1
