"""Times the SLOS chain / last layer for one tail width (FOCK_SLOS_TAIL) -- tuning helper, run under gpurun."""
import os, sys, json, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceval_b200.engine import FockEngine, prodnfact
from perceval_b200.circuit import random_unitary
n, m = int(sys.argv[1]), int(sys.argv[2])
eng = FockEngine.get(0)
U = eng.unitary(random_unitary(m, seed=0))
st = [1] * n + [0] * (m - n)
order = eng.slos_order(st)
N = eng.count(m, n)
bufs = [torch.empty(eng.count(m, n - 1), dtype=torch.complex128, device="cuda"), torch.empty(eng.count(m, n - 2), dtype=torch.complex128, device="cuda")]
probs = torch.empty(N, dtype=torch.float64, device="cuda")
psum = torch.zeros(1, dtype=torch.float64, device="cuda")
def chain(times=None):
    parent = torch.ones(1, dtype=torch.complex128, device="cuda")
    for k in range(1, n + 1):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        if k < n:
            nc = eng.count(m, k)
            buf = bufs[(n - 1 - k) % 2]
            parent = eng.slos_layer(m, k, U, order[k - 1], parent, child=buf[:nc])[:nc]
        else:
            psum.zero_()
            eng.slos_layer_probs(m, n, U, order[n - 1], parent, prodnfact(st), probs=probs, psum=psum)
        ev[1].record()
        if times is not None: times.append(ev)
for _ in range(2): chain()
torch.cuda.synchronize()
res = []
for _ in range(3):
    t = []; chain(t); torch.cuda.synchronize()
    res.append([a.elapsed_time(b) for a, b in t])
best = [min(r[i] for r in res) for i in range(n)]
print(json.dumps({"tail": os.environ.get("FOCK_SLOS_TAIL"), "kernel": os.environ.get("FOCK_SLOS_KERNEL"), "n": n, "m": m, "layer_ms": [round(x, 3) for x in best], "total_ms": round(sum(best), 3), "sum_p": psum.item(),
                  "last_GBs": (16 * eng.count(m, n - 1) + 8 * N) / best[-1] / 1e6}))
