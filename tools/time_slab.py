"""Times the slab-major chain against the FSArray-order chain on ONE GPU (same kernels, different layer layout) -- tuning
helper, run under gpurun:  python tools/time_slab.py 12 24"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceval_b200 import slab
from perceval_b200.circuit import random_unitary
from perceval_b200.engine import FockEngine

n, m = int(sys.argv[1]), int(sys.argv[2])
eng = FockEngine.get(0)
U = eng.unitary(random_unitary(m, seed=0))
st = [1] * n + [0] * (m - n)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


wa = torch.empty(eng.count(m, n - 1), dtype=torch.complex128, device=eng.device)
wb = torch.empty(eng.count(m, n - 2), dtype=torch.complex128, device=eng.device)
ms_rank = timed(lambda: eng.slos_probs(U, st, workspaces=(wa, wb)))
del wa, wb
torch.cuda.empty_cache()
chain = slab.engine_slab_chain(eng, [U], st)
last = []


def on_last(what):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    last.append(e)


ms_slab = timed(lambda: chain.run(on_last=on_last))
torch.cuda.synchronize()
last_ms = sum(last[i].elapsed_time(last[i + 1]) for i in range(0, len(last), 2)) / (len(last) // 2)
print(json.dumps({"n": n, "m": m, "ms_rank_order_chain": ms_rank, "ms_slab_chain": ms_slab, "slab_last_layer_ms": last_ms,
                  "k0": chain.plan.k0, "sum_p": float(chain.psum.item())}))
