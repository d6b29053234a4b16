N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus $N --steps 10 --warmup 3 --timeline > gpurun_out/r2n_slos_${N}gpu.json 2> gpurun_out/r2n_slos_${N}gpu.err
tail -n1 gpurun_out/r2n_slos_${N}gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('12/24 N=4 ms', d['ms_per_step'], 'value %.4g'%d['value'], d['nvlink'], d['spot_check'], 'e2e', d['e2e']['ms_per_step'])"
$TR bench.py --gpus $N --steps 3 --warmup 3 --workload permanents --perm-n 30 > gpurun_out/r2n_perm30_${N}gpu.json 2>/dev/null; tail -n1 gpurun_out/r2n_perm30_${N}gpu.json | cut -c1-200
