"""Runs one Clifford & Clifford launch (20 photons / 400 modes) for an ncu capture of cc2017_kernel -- profiling helper."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceval_b200.engine import FockEngine
from perceval_b200.circuit import random_unitary
count = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m, n = 400, 20
eng = FockEngine.get(0)
U = eng.unitary(random_unitary(m, seed=0))
smp = eng.cc2017_samples(U, [1] * n + [0] * (m - n), count, seed=0)
torch.cuda.synchronize()
print("photons per sample ok", bool((smp.sum(dim=1) == n).all()))
