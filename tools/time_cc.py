"""Times Clifford & Clifford sampling at 20 photons / 400 modes (samples/s) -- tuning helper, run under gpurun."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perceval_b200.engine import FockEngine
from perceval_b200.circuit import random_unitary
count = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m, n = 400, 20
eng = FockEngine.get(0)
U = eng.unitary(random_unitary(m, seed=0))
st = [1] * n + [0] * (m - n)
buf = torch.empty((count, m), dtype=torch.uint8, device="cuda")
eng.cc2017_samples(U, st, 2048, seed=0, out=buf[:2048])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.cc2017_samples(U, st, count, seed=0, out=buf)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(json.dumps({"lib": os.environ.get("FOCK_B200_LIB", "default"), "samples": count, "ms": ms, "samples_per_s": count / ms * 1e3,
                  "checksum": int(buf.sum(dtype=torch.int64).item())}))
