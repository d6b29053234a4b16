#!/usr/bin/env python
"""bench.py -- headline benchmark of the Fock-amplitude hot path (BASELINE.json metric).

One "step" = one full SLOS output distribution for 12 photons in 24 modes (834 451 800 states, Haar-random unitary
restated from perceval/utils/matrix.py:141-173, seed 0, input |1^12,0^12>): all 12 layers + the fused probability
epilogue, inputs resident in HBM.  `value` = states / second (whole job).  `e2e` = same through the host-buffer API
(U uploaded from pinned host memory, probabilities read back to pinned host memory inside the timed region).
Extra single-GPU figures (n=30 / n=24 Glynn permanents per second, Clifford&Clifford samples per second, measured FP64
peak) ride along in `extra`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N bench.py --gpus N ...      (N > 1)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PHOTONS, N_MODES = 12, 24
CPU_SAMPLE = (11, 22)  # bounded CPU sample: full SLOS distribution, 11 photons / 22 modes (129 024 480 states)
METRIC = "slos_amplitudes_per_s"
UNIT = "amplitudes/s"


def workload_config(extra=None):
    cfg = {"workload": f"SLOS full output distribution, {N_PHOTONS} photons / {N_MODES} modes, Haar-random unitary seed 0, "
                       f"input |1^{N_PHOTONS},0^{N_MODES - N_PHOTONS}>",
           "states": None, "l2_policy": "inputs larger than L2 (layer 11 = 4.58 GB, layer 12 probs = 6.68 GB)"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_slos_sample(steps: int = 1):
    """Times the CPU oracle (multithreaded C restatement of the reference semantics) on the bounded sample."""
    import oracle
    n, m = CPU_SAMPLE
    u = oracle.random_unitary(m, seed=0)
    st = (1,) * n + (0,) * (m - n)
    N = oracle.count(m, n)
    best = None
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        p = oracle.slos_probs(u, st)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert abs(float(p.sum()) - 1.0) < 1e-9
    return {"value": N / best, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"full SLOS distribution {n} photons / {m} modes ({N} states), CPU restatement of reference semantics "
                      f"(exqalibur not installable offline), {best:.2f} s"}, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup and 1):
        cpu_slos_sample(1)
    times = []
    base = None
    for _ in range(args.steps):
        base, dt = cpu_slos_sample(1)
        times.append(dt)
    n, m = CPU_SAMPLE
    import oracle
    N = oracle.count(m, n)
    ms = 1e3 * sum(times) / len(times)
    value = N / (ms / 1e3)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "complex128", "data": "synthetic", "config": workload_config({"states": N, "note": "bounded CPU sample, see cpu_baseline.sample"}),
            "cpu_baseline": base, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(key):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    from perceval_b200 import dist as pdist
    from perceval_b200.circuit import random_unitary
    from perceval_b200.engine import FockEngine, prodnfact

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = FockEngine.get(local_rank)
    dev = eng.device
    n, m = args.photons, args.modes
    in_state = [1] * n + [0] * (m - n)
    u_host = torch.from_numpy(random_unitary(m, seed=0)).pin_memory()
    U = eng.unitary(u_host)
    N = eng.count(m, n)
    order = eng.slos_order(in_state)
    inf = prodnfact(in_state)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- partition: replicated / gathered intermediate layers (default) or the recompute-window chain when they do not fit
    partition = args.partition
    if partition == "auto":
        free, _tot = torch.cuda.mem_get_info(dev)
        need = 16 * (eng.count(m, n - 1) + max(eng.count(m, n - 2), 1)) + 8 * (N // world + 1)
        # measured at 12 photons / 24 modes (ms / step): layers 11.7 (4 GPUs), 10.3 (8); windowed 13.1 (4), 9.5 (8)
        partition = "layers" if (need < 0.9 * free and world < 8) else "windowed"
    if partition == "windowed":
        return run_b200_windowed(args, torch, dist, pdist, eng, U, u_host, in_state, world, rank, local_rank, barrier)

    # persistent workspaces (only two layers are ever live)
    b, e = pdist.shard_range(N, rank, world)
    probs = torch.empty(e - b, dtype=torch.float64, device=dev)
    psum = torch.zeros(1, dtype=torch.float64, device=dev)
    wa = torch.empty(eng.count(m, n - 1), dtype=torch.complex128, device=dev)
    wb = torch.empty(max(eng.count(m, n - 2), 1), dtype=torch.complex128, device=dev)
    vac = torch.ones(1, dtype=torch.complex128, device=dev)

    decisions = []

    def step(Udev, last_events=None):
        """all n layers + fused probability epilogue; intermediate layers per exchange policy"""
        psum.zero_()
        parent = vac
        decisions.clear()
        for k in range(1, n):
            nc = eng.count(m, k)
            buf = wa if (n - 1 - k) % 2 == 0 else wb
            mode = "replicate" if world == 1 else (args.exchange if args.exchange != "auto" else pdist.choose_exchange(eng.count(m, k - 1), nc, world))
            decisions.append(mode)
            if mode == "replicate":
                parent = eng.slos_layer(m, k, Udev, order[k - 1], parent, child=buf[:nc])[:nc]
            else:
                sb, se = pdist.shard_range(nc, rank, world)
                shard = eng.slos_layer(m, k, Udev, order[k - 1], parent, child_begin=sb, child_end=se)
                parent = pdist.all_gather_ragged(shard, nc)
        if last_events is not None:
            last_events[0].record()
        eng.slos_layer_probs(m, n, Udev, order[n - 1], parent, inf, probs=probs, psum=psum, child_begin=b, child_end=e)
        if last_events is not None:
            last_events[1].record()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        step(U)
    barrier()
    eng.check_status()
    if world > 1:
        dist.all_reduce(psum)
    total_p = float(psum.item())
    assert abs(total_p - 1.0) < 1e-9, f"sum(p) = {total_p}"

    # ---- timed region: device-resident inputs
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(U)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = eng.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = N / (ms * 1e-3)

    # ---- dominant kernel (last layer + fused epilogue) timed live with CUDA events on the launching stream
    kms = []
    for _ in range(args.steps):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        step(U, ev)
        torch.cuda.synchronize()
        kms.append(ev[0].elapsed_time(ev[1]))
    kernel_ms = sum(kms) / len(kms)
    clk = clocks.stop()
    alg_bytes = 16.0 * eng.count(m, n - 1) + 8.0 * (e - b)  # parent layer read once + this rank's probabilities written once
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"kernel": "SLOS last layer + fused |c|^2*prod(s!)/prod(in!) epilogue (slos_thin6_kernel, csrc/slos_thin.cu, + the small-tile classes in slos_tile_kernel, csrc/slos.cu)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes,
                "traffic": load_traffic("slos_last_layer_bytes") if world == 1 else None,   # the ncu capture is of the whole layer
                "whole_chain": {"algorithmic_bytes": sum(16.0 * (eng.count(m, k - 1) + eng.count(m, k)) for k in range(1, n))
                                + 16.0 * eng.count(m, n - 1) + 8.0 * N,
                                "note": "all layers, 16 B read+write per coefficient, last layer writes 8 B probabilities"}}
    roofline["whole_chain"]["achieved"] = roofline["whole_chain"]["algorithmic_bytes"] / (ms * 1e-3) / 1e9 if world == 1 else None

    # ---- end to end through the host-buffer API: H2D of U from pinned memory, D2H of the probabilities into pinned memory
    host_probs = torch.empty(e - b, dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        """the call a user makes with HOST buffers: U from pinned host memory, distribution into pinned host memory; the
        device->host copy of each last-layer piece overlaps the computation of the next one (FockEngine.slos_probs_to_host)"""
        Ud = torch.empty_like(U)
        Ud.copy_(u_host, non_blocking=True)
        if world == 1:
            return eng.slos_probs_to_host(Ud, in_state, host_probs, pieces=args.e2e_pieces, workspaces=(wa, wb), probs=probs)
        parent = vac
        for k in range(1, n):   # intermediate layers exactly as in step()
            nc = eng.count(m, k)
            buf = wa if (n - 1 - k) % 2 == 0 else wb
            mode = args.exchange if args.exchange != "auto" else pdist.choose_exchange(eng.count(m, k - 1), nc, world)
            if mode == "replicate":
                parent = eng.slos_layer(m, k, Ud, order[k - 1], parent, child=buf[:nc])[:nc]
            else:
                sb, se = pdist.shard_range(nc, rank, world)
                parent = pdist.all_gather_ragged(eng.slos_layer(m, k, Ud, order[k - 1], parent, child_begin=sb, child_end=se), nc)
        return eng.slos_probs_to_host(Ud, in_state, host_probs, pieces=args.e2e_pieces, child_begin=b, child_end=e, probs=probs,
                                      parent=parent)

    e2e_step()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        e2e_sum = e2e_step()
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1) / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_sum)
    e2e_ms = float(t.item())
    assert abs(float(e2e_sum.item()) - 1.0) < 1e-9, "end-to-end distribution does not sum to 1"
    if world == 1:
        assert abs(float(host_probs.sum()) - 1.0) < 1e-9, "host copy of the distribution does not sum to 1"
    e2e = {"value": N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": int(u_host.numel() * 16), "d2h_bytes_per_step": int((e - b) * 8),
           "pieces": args.e2e_pieces,
           "api": "U.copy_(pinned host U) + FockEngine.slos_probs_to_host(pinned host probabilities, per rank shard): last layer in "
                  "pieces, device->host copy of piece i on a side stream under the kernel of piece i+1"}
    del host_probs

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic",
            "config": workload_config({"states": N, "photons": n, "modes": m,
                                       "partition": "single GPU" if world == 1 else f"last layer sharded by rank range over {world} GPUs; "
                                       f"intermediate layers: {sorted(set(decisions))} (exchange={args.exchange})"}),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "sum_p": total_p}

    if rank == 0 and world == 1:
        # free the big buffers before the CPU baseline and the extras
        del probs, wa, wb
        torch.cuda.empty_cache()
        if not args.no_cpu:
            try:
                line["cpu_baseline"], _ = cpu_slos_sample(1)
            except Exception as ex:  # the oracle is only a reported baseline
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        if not args.no_extras:
            line["extra"] = extras(eng, torch)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_b200_windowed(args, torch, dist, pdist, eng, U, u_host, in_state, world, rank, local_rank, barrier):
    """The same step through dist.WindowedChain: every rank recomputes exactly the parents its range of the output layer
    needs (no exchange step, no replicated layer) -- the only partition that fits 14 photons / 28 modes on 8 x 180 GB."""
    dev = eng.device
    n, m = args.photons, args.modes
    N = eng.count(m, n)
    chain = pdist.WindowedChain(eng, in_state, sub=args.sub or None)
    b, e = chain.begin, chain.end
    for _ in range(max(args.warmup, 3)):
        chain.run(U, reduce_sum=False)
    barrier()
    eng.check_status()
    psum = chain.psum.clone()
    if world > 1:
        dist.all_reduce(psum)
    total_p = float(psum.item())
    assert abs(total_p - 1.0) < 1e-9, f"sum(p) = {total_p}"

    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    events = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        chain.run(U, reduce_sum=False, last_events=events)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = eng.launch_count() - launches0
    kernel_ms = sum(a.elapsed_time(c) for a, c, _ in events) / len(events)          # per last-layer launch
    alg_bytes = sum(x for _, _, x in events) / len(events)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"kernel": "SLOS last layer + fused probability epilogue, one sub-shard, segmented parent (csrc/slos.cu, CHECK == 2)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes, "traffic": None}

    # ---- end to end: U from pinned host memory, the rank's probabilities back to the host through a pinned ring
    ring = [torch.empty(32 << 20, dtype=torch.float64).pin_memory() for _ in range(2)]   # 2 x 256 MB
    side = torch.cuda.Stream(dev)

    def e2e_step():
        Ud = torch.empty_like(U)
        Ud.copy_(u_host, non_blocking=True)
        chain.run(Ud, reduce_sum=False)
        side.wait_stream(torch.cuda.current_stream(dev))
        evs = [None, None]
        with torch.cuda.stream(side):
            for i, off in enumerate(range(0, e - b, ring[0].numel())):
                buf = ring[i % 2]
                if evs[i % 2] is not None:
                    evs[i % 2].synchronize()      # the host consumer has released this ring slot
                k = min(buf.numel(), e - b - off)
                buf[:k].copy_(chain.probs[off:off + k], non_blocking=True)
                evs[i % 2] = torch.cuda.Event()
                evs[i % 2].record(side)
        torch.cuda.current_stream(dev).wait_stream(side)

    e2e_steps = max(1, min(args.steps, 2))
    e2e_step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1) / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    clk = clocks.stop()
    e2e = {"value": N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": int(u_host.numel() * 16), "d2h_bytes_per_step": int((e - b) * 8),
           "api": "U.copy_(pinned host U) + dist.WindowedChain.run + the rank's probabilities to the host through a 2 x 256 MB pinned ring"}
    line = {"metric": METRIC, "value": N / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic",
            "config": workload_config({"states": N, "photons": n, "modes": m,
                                       "workload": f"SLOS full output distribution, {n} photons / {m} modes, Haar-random unitary seed 0, "
                                                   f"input |1^{n},0^{m - n}>",
                                       "l2_policy": "inputs larger than L2",
                                       "partition": f"recompute-window: {world} ranks x {chain.sub} sub-shards, every rank recomputes the parents "
                                                    f"its output range needs, no exchange step; workspace {chain.bytes / 1e9:.1f} GB on rank 0"}),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "sum_p": total_p}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extras(eng, torch):
    """Single-GPU figures for the other BASELINE configs (permanents n=24 / n=30, C&C sampling 20 photons / 400 modes)."""
    from perceval_b200.circuit import random_unitary
    out = {}
    try:
        fp64 = eng.measure_peak(0)
        out["fp64_fma_peak_tflops_measured"] = fp64
        out["hbm_copy_gbs_measured_here"] = eng.measure_peak(1)
        out["hbm_read_gbs_measured_here"] = eng.measure_peak(3)
        out["l2_read_gbs_measured_here"] = eng.measure_peak(2)
    except Exception as ex:
        out["peaks_error"] = str(ex)
        fp64 = None

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    for nn, B, reps in [(24, 64, 3), (30, 8, 2)]:
        mats = torch.stack([torch.from_numpy(np.ascontiguousarray(random_unitary(2 * nn, seed=s)[:nn, :nn])) for s in range(B)]).to(eng.device)
        ms = timed(lambda: eng.permanents(mats), reps)
        flops = B * (2.0 ** (nn - 1)) * (2 * nn + 6 * (nn - 1) + 2)
        rec = {"per_s": B / (ms * 1e-3), "batch": B, "ms": ms, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12}
        if fp64:
            rec["frac_of_measured_fp64_peak"] = rec["algorithmic_tflops"] / fp64
            # the pipe issues 6n-4 DFMA/DMUL per Gray step for 8n-4 algorithmic flops
            rec["fp64_pipe_utilisation_est"] = (B * (2.0 ** (nn - 1)) * (6 * nn - 4) * 2) / (ms * 1e-3) / 1e12 / fp64
        out[f"permanents_n{nn}"] = rec
    m, n, count = 400, 20, 4096
    U = eng.unitary(random_unitary(m, seed=0))
    st = [1] * n + [0] * (m - n)
    buf = torch.empty((count, m), dtype=torch.uint8, device=eng.device)
    ms = timed(lambda: eng.cc2017_samples(U, st, count, seed=0, out=buf), 1)
    out["cc2017_n20_m400"] = {"samples_per_s": count / (ms * 1e-3), "count": count, "ms": ms}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--photons", type=int, default=N_PHOTONS)
    ap.add_argument("--modes", type=int, default=N_MODES)
    ap.add_argument("--exchange", default="auto", choices=["auto", "allgather", "replicate"])
    ap.add_argument("--e2e-pieces", type=int, default=8)
    ap.add_argument("--partition", default="auto", choices=["auto", "layers", "windowed"],
                    help="layers: replicate / all-gather the intermediate layers; windowed: recompute-window chain (no exchange)")
    ap.add_argument("--sub", type=int, default=0, help="windowed: sub-shards per rank (0 = from free memory)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
