#!/usr/bin/env python
"""Full SLOS output distribution through the recompute-window partition (perceval_b200.dist.WindowedChain) -- the path for
sizes whose layers cannot be replicated, e.g. BASELINE config 5: 14 photons / 28 modes on 8 x B200.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools_run_windowed.py --photons 14 --modes 28 --steps 2

Every rank owns a contiguous range of the output layer and recomputes exactly the parents it needs (no exchange step).
Checks: sum(p) over all ranks = 1 within 1e-12; `--spot` random output states per rank are recomputed as Glynn
permanents (the Naive path, an independent kernel) and must agree to 1e-10 relative.  Prints one JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--photons", type=int, default=14)
    ap.add_argument("--modes", type=int, default=28)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sub", type=int, default=0, help="sub-shards per rank (0 = from free memory)")
    ap.add_argument("--spot", type=int, default=64)
    ap.add_argument("--mem-fraction", type=float, default=0.85)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from perceval_b200 import dist as pdist
    from perceval_b200.circuit import random_unitary
    from perceval_b200.engine import FockEngine, prodnfact

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = FockEngine.get(local_rank)
    n, m = args.photons, args.modes
    st = [1] * n + [0] * (m - n)
    u_np = random_unitary(m, seed=0)
    U = eng.unitary(u_np)
    N = eng.count(m, n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    chain = pdist.WindowedChain(eng, st, sub=args.sub or None, mem_fraction=args.mem_fraction)
    plan_s = time.time() - t0
    for _ in range(args.warmup):
        chain.run(U)
    barrier()
    eng.check_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    events = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        probs, (b, e), psum = chain.run(U, last_events=events)
    e1.record()
    barrier()
    eng.check_status()
    ms = e0.elapsed_time(e1) / args.steps
    last_ms = sum(a.elapsed_time(c) for a, c, _ in events) / args.steps
    last_bytes = sum(x for _, _, x in events) / args.steps
    t = torch.tensor([ms, last_ms, chain.bytes / 1e9], dtype=torch.float64, device=eng.device)
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    total_p = float(psum.item())

    # spot check against the Naive path: p(s) = |Perm(U_st)|^2 / (prod s! prod in!)   (reference _naive.py:46-68)
    worst = 0.0
    if args.spot > 0:
        g = torch.Generator(device="cpu").manual_seed(1234 + rank)
        idx = torch.randint(b, e, (args.spot,), generator=g, dtype=torch.int64)
        amps = eng.naive_amplitudes(U, st, out_ranks=idx.to(eng.device))
        ref = (amps.real ** 2 + amps.imag ** 2)
        got = probs[(idx - b).to(eng.device)]
        worst = float(((got - ref).abs() / ref.abs().clamp_min(1e-300)).max().item())
    w = torch.tensor([worst], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"metric": "slos_amplitudes_per_s", "value": N / (float(t[0]) * 1e-3), "unit": "amplitudes/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[0]), "ms_per_step_fastest_rank": float(tmin[0]),
                "higher_is_better": True, "scaling": "strong", "dtype": "complex128", "data": "synthetic",
                "config": {"workload": f"SLOS full output distribution, {n} photons / {m} modes, Haar-random unitary seed 0, "
                                       f"input |1^{n},0^{m - n}>", "states": N,
                           "partition": f"recompute-window, {world} ranks x {chain.sub} sub-shards, no exchange step"},
                "last_layer_ms_slowest_rank": float(t[1]), "last_layer_algorithmic_bytes_rank0": last_bytes,
                "workspace_GB_max_rank": float(t[2]), "plan_s_rank0": plan_s, "sum_p": total_p,
                "sum_p_error": abs(total_p - 1.0), "spot_check": {"states_per_rank": args.spot, "against": "Glynn permanents (naive_amplitudes)",
                                                                  "worst_rel_err": float(w[0])}}
        print(json.dumps(line), flush=True)
    assert abs(total_p - 1.0) < 1e-9, total_p
    assert float(w[0]) < 1e-9, float(w[0])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
