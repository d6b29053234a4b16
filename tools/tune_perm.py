"""Times batched Glynn permanents (tuning helper, run under gpurun)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceval_b200.engine import FockEngine
from perceval_b200.circuit import random_unitary
eng = FockEngine.get(0)
fp64 = eng.measure_peak(0)
res = {"fp64_peak": fp64, "waves": os.environ.get("FOCK_GLYNN_WAVES")}
for n, B, reps in [(24, 64, 3), (30, 8, 2), (32, 4, 1)]:
    mats = torch.stack([torch.from_numpy(np.ascontiguousarray(random_unitary(2 * n, seed=s)[:n, :n])) for s in range(B)]).cuda()
    eng.permanents(mats); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = eng.permanents(mats)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    fl = B * 2.0 ** (n - 1) * (8 * n - 4)
    res[f"n{n}"] = {"per_s": round(B / ms * 1e3, 2), "ms": round(ms, 2), "tflops": round(fl / ms / 1e9, 2), "frac": round(fl / ms / 1e9 / fp64, 3)}
print(json.dumps(res))
