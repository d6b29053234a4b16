"""NVLink point-to-point bandwidth as the slab exchange sees it (NCCL send / recv through torch.distributed) -- run under
torchrun with >= 2 ranks:  torchrun --nproc-per-node 2 tools/p2p_bw.py"""
import json, os
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
out = {}
for mb in (64, 1024):
    n = mb * (1 << 20) // 8
    a = torch.ones(n, dtype=torch.float64, device="cuda")
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    for name in ("ring_one_way", "ring_both_ways"):
        def go():
            nxt, prv = (rank + 1) % world, (rank - 1) % world
            ops = [dist.P2POp(dist.irecv, b, prv), dist.P2POp(dist.isend, a, nxt)]
            if name == "ring_both_ways":
                ops += [dist.P2POp(dist.irecv, b2, nxt), dist.P2POp(dist.isend, a, prv)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        b2 = torch.empty(n, dtype=torch.float64, device="cuda") if name == "ring_both_ways" else None
        for _ in range(3):
            go()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            go()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f"{name}_{mb}MB_GBs_per_direction"] = round(mb / 1024 * 1.073741824 / (ms * 1e-3), 1)
if rank == 0:
    print(json.dumps({"world": world, "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}, **out}))
dist.destroy_process_group()
