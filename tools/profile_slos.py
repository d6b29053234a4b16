"""Runs one SLOS chain (for ncu capture of the last-layer kernel) -- profiling helper, run under gpurun + ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceval_b200.engine import FockEngine
from perceval_b200.circuit import random_unitary
n, m = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = FockEngine.get(0)
U = eng.unitary(random_unitary(m, seed=0))
st = [1] * n + [0] * (m - n)
for _ in range(reps):
    probs, psum, _ = eng.slos_probs(U, st)
torch.cuda.synchronize()
print("sum_p", psum.item())
