"""Runs one Glynn batch and one Clifford&Clifford batch (for ncu capture) -- profiling helper, run under gpurun + ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceval_b200.engine import FockEngine
from perceval_b200.circuit import random_unitary
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = FockEngine.get(0)
mats = torch.stack([torch.from_numpy(np.ascontiguousarray(random_unitary(2 * n, seed=s)[:n, :n])) for s in range(B)]).cuda()
out = eng.permanents(mats)
torch.cuda.synchronize()
print("perm0", complex(out[0].item()))
if len(sys.argv) > 3:
    m, nn, cnt = 400, 20, int(sys.argv[3])
    U = eng.unitary(random_unitary(m, seed=0))
    s = eng.cc2017_samples(U, [1] * nn + [0] * (m - nn), cnt, seed=0)
    torch.cuda.synchronize()
    print("samples", s.shape)
