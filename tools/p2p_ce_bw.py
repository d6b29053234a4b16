"""NVLink peer copies through the copy engines (cudaMemcpyPeerAsync = tensor.copy_ across devices), one process, two GPUs:
one direction and both directions at once -- the transport a pull-style halo exchange would use instead of NCCL send/recv."""
import json
import torch

assert torch.cuda.device_count() >= 2
out = {}
for mb in (64, 1024):
    n = mb * (1 << 20) // 8
    a0 = torch.ones(n, dtype=torch.float64, device="cuda:0"); b0 = torch.empty(n, dtype=torch.float64, device="cuda:0")
    a1 = torch.ones(n, dtype=torch.float64, device="cuda:1"); b1 = torch.empty(n, dtype=torch.float64, device="cuda:1")
    s0, s1 = torch.cuda.Stream(device="cuda:0"), torch.cuda.Stream(device="cuda:1")
    for mode in ("one_way", "both_ways"):
        def go():
            with torch.cuda.stream(s1):
                b1.copy_(a0, non_blocking=True)          # pull 0 -> 1, issued on device 1's stream
            if mode == "both_ways":
                with torch.cuda.stream(s0):
                    b0.copy_(a1, non_blocking=True)      # pull 1 -> 0
        for _ in range(3):
            go()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s1):
            e0.record()
        for _ in range(5):
            go()
        with torch.cuda.stream(s1):
            e1.record()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        ms = e0.elapsed_time(e1) / 5
        out[f"{mode}_{mb}MB_GBs_per_direction"] = round(mb / 1024 * 1.073741824 / (ms * 1e-3), 1)
print(json.dumps({"transport": "copy engine peer copy", **out}))
