N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus $N --steps 3 --warmup 1 --photons 14 --modes 28 --timeline > gpurun_out/r2m_slos_14_28_${N}gpu.json 2> gpurun_out/r2m_slos_14_28_${N}gpu.err
tail -n1 gpurun_out/r2m_slos_14_28_${N}gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('14/28 ms', d['ms_per_step'], 'value %.4g'%d['value'], d['nvlink'], 'agg', d['roofline']['aggregate']['frac_of_n_gpus_x_peak'], d['spot_check'], d['sum_p'])
for r,t in enumerate(d['timeline_ms_per_rank']): print(r, [x for x in t if x[0].startswith('layer1')][-3:], round(sum(x[1] for x in t),1))
"; grep -m3 Error gpurun_out/r2m_slos_14_28_${N}gpu.err
