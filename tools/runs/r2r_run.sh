TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_slos_2gpu.json 2> gpurun_out/r2r_slos_2gpu.err
tail -n1 gpurun_out/r2r_slos_2gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 ms', d['ms_per_step'], d['config']['partition_name'], d['nvlink'], d['spot_check'], 'e2e', d['e2e']['ms_per_step'], 'sum_p', d['sum_p'])"
grep -m3 "Error" gpurun_out/r2r_slos_2gpu.err
