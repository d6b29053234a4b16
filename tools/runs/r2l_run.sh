N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
show() { tail -n1 "$1" | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$2', 'ms', round(d['ms_per_step'],3), 'value', '%.4g'%d['value'], 'kernel', d['roofline'].get('kernel_ms'), d.get('nvlink'), 'e2e', round(d['e2e']['ms_per_step'],1), 'frac', d['roofline'].get('frac'), 'agg', d['roofline'].get('aggregate',{}).get('frac_of_n_gpus_x_peak'), d.get('spot_check',{}).get('worst_rel_err'))
    for r,t in enumerate(d.get('timeline_ms_per_rank',[])):
        if r in (0,3,4,7): print('   ', r, t)
except Exception as e: print('$2', 'no json', e)"; }
$TR bench.py --gpus $N --steps 10 --warmup 3 --timeline > gpurun_out/r2l_slos_${N}gpu.json 2> gpurun_out/r2l_slos_${N}gpu.err; show gpurun_out/r2l_slos_${N}gpu.json slos_default; grep -m3 "Error" gpurun_out/r2l_slos_${N}gpu.err
if [ "$N" = "8" ]; then
  $TR bench.py --gpus $N --steps 2 --warmup 1 --photons 14 --modes 28 --timeline > gpurun_out/r2l_slos_14_28_${N}gpu.json 2> gpurun_out/r2l_slos_14_28_${N}gpu.err; show gpurun_out/r2l_slos_14_28_${N}gpu.json slos_14_28; grep -m3 "Error" gpurun_out/r2l_slos_14_28_${N}gpu.err
fi
for pn in 30 32; do
  $TR bench.py --gpus $N --steps 3 --warmup 3 --workload permanents --perm-n $pn > gpurun_out/r2l_perm${pn}_${N}gpu.json 2> gpurun_out/r2l_perm${pn}_${N}gpu.err; show gpurun_out/r2l_perm${pn}_${N}gpu.json perm$pn
done
$TR bench.py --gpus $N --steps 3 --warmup 3 --workload cc2017 > gpurun_out/r2l_cc_${N}gpu.json 2> gpurun_out/r2l_cc_${N}gpu.err; show gpurun_out/r2l_cc_${N}gpu.json cc2017
