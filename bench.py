#!/usr/bin/env python
"""bench.py -- headline benchmark of the Fock-amplitude hot path (BASELINE.json metric).

Default workload (`--workload slos`): one "step" = one full SLOS output distribution for 12 photons in 24 modes
(834 451 800 states, Haar-random unitary restated from perceval/utils/matrix.py:141-173, seed 0, input |1^12,0^12>): all 12
layers + the fused probability epilogue.  On one GPU the step is timed THROUGH THE BACKEND API
(`BackendFactory.get_backend("SLOS_B200")` -> set_circuit -> set_input_state -> all_prob_tensor): `value` with the unitary
resident in HBM, `e2e` with the unitary in pinned host memory and the distribution written to pinned host memory
(`all_prob_into`).  On N > 1 GPUs (torchrun, one process per GPU) the same chain runs under one of the partitions of
perceval_b200/dist.py.  `--workload permanents` / `--workload cc2017` put the other BASELINE configs (n = 30 / 32 Glynn
permanents, Clifford & Clifford sampling at 20 photons / 400 modes) on the same 1/2/4/8-GPU footing.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload slos|permanents|cc2017]
    torchrun --nnodes=1 --nproc-per-node N bench.py --gpus N ...      (N > 1)
"""
from __future__ import annotations

import argparse
import hashlib
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


def slos_workload(n: int, m: int) -> str:
    return f"SLOS full output distribution, {n} photons / {m} modes, Haar-random unitary seed 0, input |1^{n},0^{m - n}>"


def slos_config(n: int, m: int, states: int, **extra) -> dict:
    cfg = {"workload": slos_workload(n, m), "states": states, "photons": n, "modes": m,
           "l2_policy": "inputs larger than L2 (12/24: layer 11 = 4.58 GB, probabilities = 6.68 GB)"}
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
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must be given the box's cores explicitly, before the
    OpenMP runtime of the oracle is loaded."""
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    os.environ.pop("OMP_THREAD_LIMIT", None)


def cpu_slos_once(n: int, m: int):
    import oracle
    u = oracle.random_unitary(m, seed=0)
    st = (1,) * n + (0,) * (m - n)
    t0 = time.perf_counter()
    p = oracle.slos_probs(u, st)
    dt = time.perf_counter() - t0
    assert abs(float(p.sum()) - 1.0) < 1e-9
    return dt, int(p.shape[0])


def cpu_slos_sample():
    """cpu_baseline of the main arm: the CPU oracle (multithreaded C restatement of the reference semantics) on a bounded
    sample of the workload, rank 0, N = 1 only."""
    use_all_host_threads()
    import oracle
    n, m = CPU_SAMPLE
    dt, N = cpu_slos_once(n, m)
    return {"value": N / dt, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"full SLOS distribution {n} photons / {m} modes ({N} states), CPU restatement of reference semantics "
                      f"(exqalibur not installable offline), {dt:.2f} s"}


def run_reference(args):
    """The reference arm: the CPU implementation of the path (oracle port, all host threads) on the same metric.  SLOS: the
    real 12 photons / 24 modes when host memory and the time budget allow (25 GB, ~7 x the 11/22 probe per step), otherwise
    the bounded 11 / 22 sample -- labelled as what it is in `config`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    import oracle
    cores = oracle.num_threads()
    if args.workload == "permanents":
        return run_reference_permanents(args, oracle, cores)
    if args.workload == "cc2017":
        return run_reference_cc2017(args, oracle, cores)
    n, m = args.photons, args.modes
    probe_dt, _ = cpu_slos_once(*CPU_SAMPLE)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    N_full = oracle.count(m, n)
    need = 16 * (N_full + oracle.count(m, n - 1)) + 8 * N_full + (2 << 30)
    ratio = N_full * m / (oracle.count(CPU_SAMPLE[1], CPU_SAMPLE[0]) * CPU_SAMPLE[1])
    predicted = probe_dt * ratio * (args.steps + args.warmup)
    full = avail > need and predicted < args.reference_budget_s
    if not full:
        n, m = CPU_SAMPLE
    for _ in range(args.warmup):
        cpu_slos_once(n, m)
    times = []
    for _ in range(args.steps):
        dt, N = cpu_slos_once(n, m)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = N / (ms / 1e3)
    sample = (f"full SLOS distribution {n} photons / {m} modes ({N} states) per step, CPU restatement of reference semantics "
              f"(exqalibur not installable offline), OpenMP over {cores} threads")
    if not full:
        sample += (f"; the named 12/24 workload was NOT run (needs {need / 1e9:.0f} GB host memory, {avail / 1e9:.0f} GB available; "
                   f"predicted {predicted:.0f} s for {args.steps}+{args.warmup} steps against a budget of {args.reference_budget_s} s)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "complex128", "data": "synthetic", "config": slos_config(n, m, N),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def haar_submatrices(nn: int, B: int):
    from perceval_b200.circuit import random_unitary
    return np.stack([np.ascontiguousarray(random_unitary(2 * nn, seed=s)[:nn, :nn]) for s in range(B)])


def perm_flops(nn: int) -> float:
    return (2.0 ** (nn - 1)) * (2 * nn + 6 * (nn - 1) + 2)


def run_reference_permanents(args, oracle, cores):
    nn, B = args.perm_n, 1
    mats = haar_submatrices(nn, B)
    G = 1 << (nn - 1)
    frac = 1
    probe0 = time.perf_counter()
    oracle.permanent(mats[0], 0, G >> 6)
    per_full = (time.perf_counter() - probe0) * 64
    while per_full / frac * (args.steps + args.warmup) > args.reference_budget_s and frac < 4096:
        frac *= 2
    g1 = G // frac
    for _ in range(args.warmup):
        oracle.permanent(mats[0], 0, g1)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle.permanent(mats[0], 0, g1)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = (1.0 / frac) / (ms / 1e3)
    sample = f"Gray codes [0, 2^{nn - 1}/{frac}) of one n = {nn} Glynn permanent per step, OpenMP over {cores} threads (CPU restatement of xq.permanent_cx)"
    line = {"impl": "reference", "metric": "glynn_permanents_per_s", "value": value, "unit": "permanents/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "complex128", "data": "synthetic", "config": perm_config(nn, args.batch),
            "cpu_baseline": {"value": value, "unit": "permanents/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "permanents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference_cc2017(args, oracle, cores):
    m, n = args.modes if args.modes != N_MODES else 400, args.photons if args.photons != N_PHOTONS else 20
    u = oracle.random_unitary(m, seed=0)
    st = (1,) * n + (0,) * (m - n)
    count = 4 * cores
    t0 = time.perf_counter()
    oracle.cc2017_samples(u, st, count, seed=0)
    per = (time.perf_counter() - t0) / count
    count = max(cores, min(args.samples, int(args.reference_budget_s / max(per, 1e-9) / (args.steps + args.warmup))))
    for _ in range(args.warmup):
        oracle.cc2017_samples(u, st, count, seed=0)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle.cc2017_samples(u, st, count, seed=0)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = count / (ms / 1e3)
    sample = f"{count} samples per step at {n} photons / {m} modes, OpenMP over {cores} threads (CPU restatement of xq.Clifford2017)"
    line = {"impl": "reference", "metric": "cc2017_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "complex128", "data": "synthetic", "config": cc_config(n, m, args.samples),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm: helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash() -> str:
    h = hashlib.sha256()
    for rel in ["perceval_b200/csrc/slos.cu", "perceval_b200/csrc/slos_thin.cu", "perceval_b200/csrc/slos_tile.cuh",
                "perceval_b200/csrc/slos_mu.cu"]:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def load_traffic():
    """DRAM bytes of the dominant launch from the committed ncu capture (profiles/ncu_traffic.json, written by
    tools/ncu_traffic.py) -- only if that capture is of the kernel sources as they are now."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None, "no ncu capture committed"
    if d.get("source_sha") != kernel_source_hash():
        return None, f"stale: {d.get('source_csv')} was captured from other kernel sources"
    return d.get("slos_last_layer_bytes"), d.get("source_csv")


def dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL send / recv between two B200s: 634 GB/s one way with the default channel count, 672 GB/s with 64 (tools/p2p_bw.py)
        for key in ("NCCL_MAX_NCHANNELS", "NCCL_MIN_NCHANNELS", "NCCL_MIN_P2P_NCHANNELS", "NCCL_MAX_P2P_NCHANNELS"):
            os.environ.setdefault(key, "64")
        try:    # NCCL's copy kernels on a high-priority stream: a halo transfer must not queue behind a 130 000-CTA compute grid
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
        except Exception:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rmax(x: float) -> float:
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return torch, dist, world, rank, local_rank, barrier, rmax


def timed_steps(torch, barrier, rmax, fn, steps: int):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return rmax(e0.elapsed_time(e1) / steps)


def spot_check(torch, eng, u_np, in_state, probs, begin, end, count, seed):
    """`count` random outputs of this rank's probabilities against the CPU oracle's permanent-based amplitude
    (oracle.naive_amplitude, reference _naive.py:46-68) -- outside every timed region."""
    import oracle
    if count <= 0 or end <= begin:
        return 0.0
    m, n = len(in_state), sum(in_state)
    rng = np.random.default_rng(seed)
    ranks = rng.integers(begin, end, count)
    states = oracle.unrank_batch(m, n, ranks.astype(np.uint64))
    got = probs[torch.from_numpy(ranks - begin).to(probs.device)].cpu().numpy()
    ref = np.array([abs(oracle.naive_amplitude(u_np, tuple(in_state), tuple(int(x) for x in s))) ** 2 for s in states])
    return float(np.abs(got - ref).max() / max(ref.max(), 1e-300))


def spot_check_slab(torch, chain, u_np, in_state, count, seed):
    """the same check for the slab partition, whose output is stored compactly in slab-major order (chain.out_slices)"""
    import oracle
    from perceval_b200 import partition as P
    if count <= 0 or chain.probs.numel() == 0:
        return 0.0
    L = chain.plan.layout
    n = chain.n
    rng = np.random.default_rng(seed)
    pos = rng.integers(0, chain.probs.numel(), count)
    got = chain.probs[torch.from_numpy(pos).to(chain.probs.device)].cpu().numpy()
    ref = []
    for x in pos:
        for w, a, b, off, ln in chain.out_slices:
            if off <= x < off + ln:
                S = L.S[n][w]
                rho, t = a + (x - off) // S, (x - off) % S
                state = P.unrank(L.p, w, int(rho)) + P.unrank(L.D, n - w, int(t))
                ref.append(abs(oracle.naive_amplitude(u_np, tuple(in_state), tuple(state))) ** 2)
                break
    ref = np.array(ref)
    return float(np.abs(got - ref).max() / max(ref.max(), 1e-300))


# ------------------------------------------------------------------------------------------------ SLOS, one GPU: through the backend
def run_slos_single(args, torch, rank, local_rank, barrier, rmax):
    import perceval_b200 as pb
    from perceval_b200.circuit import random_unitary
    from perceval_b200.engine import FockEngine

    n, m = args.photons, args.modes
    st = pb.BasicState([1] * n + [0] * (m - n))
    u_np = random_unitary(m, seed=0)
    u_host = torch.from_numpy(u_np).pin_memory()
    eng = FockEngine.get(local_rank)
    N = eng.count(m, n)
    backend = pb.BackendFactory.get_backend("SLOS_B200", device=local_rank)
    circ_dev = pb.UnitaryCircuit(eng.unitary(u_host))     # unitary resident in HBM: set_circuit copies nothing
    circ_host = pb.UnitaryCircuit(u_host)                 # unitary in pinned host memory (end-to-end leg)
    out = {}

    def step():
        """the calls a user makes: a new circuit of the same size invalidates the deployed input, set_input_state +
        all_prob_tensor recompute the whole chain (all n layers + fused probability epilogue)"""
        backend.set_circuit(circ_dev)
        backend.set_input_state(st)
        out["probs"] = backend.all_prob_tensor()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    res = backend._results[st]
    total_p = float(res.psum.item())
    assert abs(total_p - 1.0) < 1e-9, f"sum(p) = {total_p}"
    worst = spot_check(torch, eng, u_np, [int(x) for x in st], out["probs"], 0, N, args.spot, 1234) if args.spot else None
    if worst is not None:
        assert worst < 1e-9, worst

    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count()
    layers0 = backend.stats["layers_computed"]
    ms = timed_steps(torch, barrier, rmax, step, args.steps)
    launches = eng.launch_count() - launches0
    assert backend.stats["layers_computed"] - layers0 == n * args.steps, "a step did not recompute the chain"
    value = N / (ms * 1e-3)

    # dominant kernel (last layer + fused epilogue): CUDA events recorded by the library around that launch, on its stream
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    eng.profile_events(*ev)
    kms = []
    for _ in range(args.steps):
        step()
        torch.cuda.synchronize()
        kms.append(ev[0].elapsed_time(ev[1]))
    eng.profile_events(None, None)
    kernel_ms = sum(kms) / len(kms)
    clk = clocks.stop()
    alg_bytes = 16.0 * eng.count(m, n - 1) + 8.0 * N
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = load_traffic() if (n, m) == (N_PHOTONS, N_MODES) else (None, "capture is of 12 photons / 24 modes")
    chain_bytes = sum(16.0 * (eng.count(m, k - 1) + eng.count(m, k)) for k in range(1, n)) + alg_bytes
    roofline = {"kernel": "SLOS last layer + fused |c|^2*prod(s!)/prod(in!) epilogue (slos_thin6_kernel, csrc/slos_thin.cu, + the small-tile "
                          "classes in slos_tile_kernel, csrc/slos.cu)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes, "traffic": traffic, "traffic_source": traffic_src,
                "whole_chain": {"algorithmic_bytes": chain_bytes, "achieved": chain_bytes / (ms * 1e-3) / 1e9,
                                "frac": chain_bytes / (ms * 1e-3) / 1e9 / peak,
                                "note": "all layers, 16 B read + write per coefficient, last layer writes 8 B probabilities"}}

    # end to end through the same API with HOST buffers
    backend._drop_results()
    out.clear()
    torch.cuda.empty_cache()
    host_probs = torch.empty(N, dtype=torch.float64).pin_memory()
    sums = []

    def e2e_step():
        backend.set_circuit(circ_host)          # H2D of the unitary from pinned memory
        backend.set_input_state(st)             # large input: deferred, so that ...
        sums.append(backend.all_prob_into(host_probs, pieces=args.e2e_pieces))   # ... the last layer overlaps the D2H copies

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    e2e_ms = timed_steps(torch, barrier, rmax, e2e_step, e2e_steps)
    assert all(abs(s_ - 1.0) < 1e-9 for s_ in sums), sums
    assert abs(float(host_probs.sum()) - 1.0) < 1e-9, "host copy of the distribution does not sum to 1"
    e2e = {"value": N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": int(u_host.numel() * 16), "d2h_bytes_per_step": int(N * 8), "pieces": args.e2e_pieces,
           "api": "SLOSB200Backend.set_circuit(UnitaryCircuit(pinned host U)) + set_input_state + all_prob_into(pinned host float64): "
                  "last layer in pieces, device->host copy of piece i on a side stream under the kernel of piece i+1"}
    del host_probs

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
            "config": slos_config(n, m, N, partition="single GPU",
                                  api='BackendFactory.get_backend("SLOS_B200"): set_circuit + set_input_state + all_prob_tensor per step'),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "sum_p": total_p,
            "spot_check": {"outputs": args.spot, "against": "oracle.naive_amplitude (CPU permanents)", "worst_rel_err": worst}}
    backend._drop_results()
    torch.cuda.empty_cache()
    if not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_slos_sample()
        except Exception as ex:  # the oracle is only a reported baseline
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": f"failed: {ex}"}
    if not args.no_extras:
        line["extra"] = extras(eng, torch)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ SLOS, N > 1 GPUs
def run_slos_multi(args, torch, dist, world, rank, local_rank, barrier, rmax):
    from perceval_b200 import dist as pdist
    from perceval_b200.circuit import random_unitary
    from perceval_b200.engine import FockEngine

    eng = FockEngine.get(local_rank)
    dev = eng.device
    n, m = args.photons, args.modes
    in_state = [1] * n + [0] * (m - n)
    u_np = random_unitary(m, seed=0)
    u_host = torch.from_numpy(u_np).pin_memory()
    U = eng.unitary(u_host)
    N = eng.count(m, n)
    partition = args.partition
    if partition == "auto":
        from perceval_b200 import slab as pslab
        free, _tot = torch.cuda.mem_get_info(dev)
        sp = pslab.SlabPlan(m, n, world, shard_min=args.shard_min, pieces=args.pieces or (4 if pslab.SlabLayout(m, n).p >= 10 else 1))
        need = max(16 * sum(sp.buffer_elems(q)) + 8 * sp.own_elems(n, q) for q in range(world)) + pdist.tail_table_bytes(n)
        partition = "slab" if need < 0.85 * free else "windowed"
    events = []          # (begin, end) CUDA events around every last-layer launch of a step
    alg_bytes = [0.0]

    if partition == "windowed":
        chain = pdist.WindowedChain(eng, in_state, sub=args.sub or None)
        b, e = chain.begin, chain.end
        get_probs = lambda: chain.probs
        get_sum = lambda: chain.psum
        last = []

        def step(Udev=U):
            last.clear()
            chain.run(Udev, reduce_sum=False, last_events=last)
            events.extend((a, c) for a, c, _ in last)
            alg_bytes[0] = sum(x for _, _, x in last)
        desc = (f"recompute-window: {world} ranks x {chain.sub} sub-shards, every rank recomputes the parents its output range needs, no "
                f"exchange step; workspace {chain.bytes / 1e9:.1f} GB on rank 0")
        nvlink = {"bytes_received_per_step": 0, "bytes_sent_per_step": 0}
    elif partition == "slab":
        from perceval_b200 import slab as pslab
        U_ref = [U]
        pieces = args.pieces or (4 if pslab.SlabLayout(m, n).p >= 10 else 1)
        chain = pslab.engine_slab_chain(eng, U_ref, in_state, shard_min=args.shard_min, pieces=pieces)
        plan = chain.plan
        b, e = 0, chain.probs.numel()          # compact slab-major storage of this rank's prefixes (chain.out_slices)
        get_probs = lambda: chain.probs
        get_sum = lambda: chain.psum

        def on_last(what):
            evt = torch.cuda.Event(enable_timing=True)
            evt.record()
            if what == "begin":
                events.append([evt, None])
            else:
                events[-1][1] = evt

        marks = []

        def mark(label):
            evt = torch.cuda.Event(enable_timing=True)
            evt.record()
            marks.append((label, evt))

        def step(Udev=U):
            U_ref[0] = Udev
            marks.clear()
            if args.timeline:
                mark("start")
            chain.run(reduce_sum=False, on_last=on_last, mark=mark if args.timeline else None)
        Lh = plan.layout
        rows_last = sum((hi - lo) * Lh.S[n - 1][w - 1] for w, segs in plan.rows[rank].items() for lo, hi in segs)
        tails_last = sum((bb - aa) * Lh.S[n - 1][w] for w, aa, bb in plan.own[rank] if w <= n - 1)
        alg_bytes[0] = 16.0 * (rows_last + tails_last) + 8.0 * (e - b)
        desc = (f"slab partition: layers < {plan.k0} replicated; layers {plan.k0}..{n} stored slab-major (prefix weight over the first "
                f"{Lh.p} modes, prefix rank, tail rank), every rank owns a fixed run of prefixes, tail parents local, prefix rows "
                f"received over NVLink (NCCL send/recv of contiguous slab slices, {pieces} piece(s) / exchange group(s) per layer); output "
                f"stays sharded in slab-major order")
        nvlink = {"bytes_received_per_step": int(chain.bytes_received), "bytes_sent_per_step": int(chain.bytes_sent)}
    else:
        U_ref = [U]
        shard_min = (1 << 62) if partition == "replicate" else args.shard_min
        chain = pdist.engine_exchange_chain(eng, U_ref, in_state, pieces=args.pieces or 4, shard_min=shard_min)
        b, e = chain.begin, chain.end
        get_probs = lambda: chain.probs
        get_sum = lambda: chain.psum

        def on_piece(j, what):
            evt = torch.cuda.Event(enable_timing=True)
            evt.record()
            if what == "begin":
                events.append([evt, None])
            else:
                events[-1][1] = evt

        def step(Udev=U):
            U_ref[0] = Udev
            chain.run(reduce_sum=False, on_last_piece=on_piece)
        plan = chain.plan
        need_last = sum(hi - lo for segs in plan.need[n][rank] for lo, hi in segs) if plan.k0 < n else eng.count(m, n - 1)
        alg_bytes[0] = 16.0 * min(need_last, eng.count(m, n - 1)) + 8.0 * (e - b)
        desc = (f"owner-computes + halo exchange: layers < {plan.k0} replicated, layers {plan.k0}..{n} cut in {world} rank ranges "
                f"x {args.pieces or 4} pieces, parents sent over NVLink (NCCL send/recv) in groups ordered by the consumer's pieces"
                if plan.k0 < n else f"output layer cut in {world} rank ranges, layers 1..{n - 1} replicated on every rank, no exchange")
        nvlink = {"bytes_received_per_step": int(chain.bytes_received), "bytes_sent_per_step": int(chain.bytes_sent)}

    warm = max(args.warmup, 3)
    for _ in range(warm):
        events.clear()
        step()
    barrier()
    eng.check_status()
    psum = get_sum().clone()
    dist.all_reduce(psum)
    total_p = float(psum.item())
    assert abs(total_p - 1.0) < 1e-9, f"sum(p) = {total_p}"
    if not args.spot:
        worst = None
    elif partition == "slab":
        worst = rmax(spot_check_slab(torch, chain, u_np, in_state, args.spot, 1234 + rank))
    else:
        worst = rmax(spot_check(torch, eng, u_np, in_state, get_probs(), b, e, args.spot, 1234 + rank))
    if worst is not None:
        assert worst < 1e-9, worst

    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count()
    events.clear()
    ms = timed_steps(torch, barrier, rmax, step, args.steps)
    launches = eng.launch_count() - launches0
    torch.cuda.synchronize()
    kernel_ms = sum(a.elapsed_time(c) for a, c in events) / args.steps
    events.clear()
    kernel_ms_max = rmax(kernel_ms)
    peak, peak_src = load_peaks()
    achieved = alg_bytes[0] / (kernel_ms * 1e-3) / 1e9
    roofline = {"kernel": "SLOS last layer + fused probability epilogue, this rank's range (sum of its launches per step)",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "kernel_ms": kernel_ms, "kernel_ms_slowest_rank": kernel_ms_max, "algorithmic_bytes": alg_bytes[0], "traffic": None,
                "aggregate": {"algorithmic_bytes_whole_chain": sum(16.0 * (eng.count(m, k - 1) + eng.count(m, k)) for k in range(1, n))
                              + 16.0 * eng.count(m, n - 1) + 8.0 * N}}
    roofline["aggregate"]["frac_of_n_gpus_x_peak"] = roofline["aggregate"]["algorithmic_bytes_whole_chain"] / (ms * 1e-3) / 1e9 / (world * peak)

    # end to end: U from pinned host memory, this rank's probabilities to the host through a pinned ring (2 x 256 MB: the
    # consumer side of a host pipeline; pinning the whole 35 GB shard of a 14/28 run on 8 ranks would take longer than the run)
    ring = [torch.empty(32 << 20, dtype=torch.float64).pin_memory() for _ in range(2)]
    side = torch.cuda.Stream(dev)
    host_sum = [0.0]

    def e2e_step(consume: bool = False):
        Ud = torch.empty_like(U)
        Ud.copy_(u_host, non_blocking=True)
        step(Ud)
        probs_dev = get_probs()
        side.wait_stream(torch.cuda.current_stream(dev))
        evs = [None, None]
        total = 0.0
        with torch.cuda.stream(side):
            for i, off in enumerate(range(0, probs_dev.numel(), ring[0].numel())):
                buf = ring[i % 2]
                if evs[i % 2] is not None:
                    evs[i % 2][0].synchronize()           # the previous copy into this slot has landed: the slot is free again
                    if consume:
                        total += float(buf[:evs[i % 2][1]].sum())
                k = min(buf.numel(), probs_dev.numel() - off)
                buf[:k].copy_(probs_dev[off:off + k], non_blocking=True)
                ev_ = torch.cuda.Event()
                ev_.record(side)
                evs[i % 2] = (ev_, k)
            for slot in evs:
                if slot is not None:
                    slot[0].synchronize()
                    if consume:
                        total += float(ring[evs.index(slot)][:slot[1]].sum())
        torch.cuda.current_stream(dev).wait_stream(side)
        host_sum[0] = total

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step(consume=True)      # untimed pass that also sums what arrived on the host
    consumed = host_sum[0]
    e2e_ms = timed_steps(torch, barrier, rmax, e2e_step, e2e_steps)
    clk = clocks.stop()
    hs = torch.tensor([consumed], dtype=torch.float64, device=dev)
    dist.all_reduce(hs)
    assert abs(float(hs.item()) - 1.0) < 1e-9, "host copies of the distribution do not sum to 1"
    e2e = {"value": N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": int(u_host.numel() * 16), "d2h_bytes_per_step": int((e - b) * 8),
           "api": "per rank: U.copy_(pinned host U) + the partition's chain + this rank's probabilities to the host through a 2 x 256 MB pinned ring (consumed = summed on the host)"}
    nvlink = {"bytes_received_per_step_this_rank": nvlink["bytes_received_per_step"], "bytes_sent_per_step_this_rank": nvlink["bytes_sent_per_step"],
              "bytes_received_per_step_max_rank": int(rmax(nvlink["bytes_received_per_step"])),
              "bytes_sent_per_step_max_rank": int(rmax(nvlink["bytes_sent_per_step"]))}
    line = {"metric": METRIC, "value": N / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic", "config": slos_config(n, m, N, partition=desc, partition_name=partition),
            "roofline": roofline, "nvlink": nvlink, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "sum_p": total_p,
            "spot_check": {"outputs_per_rank": args.spot, "against": "oracle.naive_amplitude (CPU permanents)", "worst_rel_err": worst}}
    if args.timeline and partition == "slab":
        step()
        torch.cuda.synchronize()
        tl = [(marks[i + 1][0], round(marks[i][1].elapsed_time(marks[i + 1][1]), 3)) for i in range(len(marks) - 1)]
        gathered = [None] * world
        dist.all_gather_object(gathered, tl)
        line["timeline_ms_per_rank"] = gathered
    if rank == 0:
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ permanents / sampling workloads
def perm_config(nn: int, B: int) -> dict:
    return {"workload": f"Naive backend: batch of {B} Glynn permanents, n = {nn}, top-left n x n blocks of Haar-random 2n x 2n unitaries "
                        f"(seeds 0..{B - 1})", "n": nn, "batch": B,
            "l2_policy": "compute-bound: matrices live in shared memory, no reuse between launches to flush"}


def cc_config(n: int, m: int, count: int) -> dict:
    return {"workload": f"Clifford & Clifford 2017 boson sampling, {n} photons / {m} modes, {count} samples, Haar-random unitary seed 0, "
                        f"input |1^{n},0^{m - n}>", "photons": n, "modes": m, "samples": count,
            "l2_policy": "compute-bound: per-sample state lives in shared memory"}


def measure_fp64(eng):
    try:
        return eng.measure_peak(0)
    except Exception:
        return None


def run_permanents(args, torch, dist, world, rank, local_rank, barrier, rmax):
    from perceval_b200 import dist as pdist
    from perceval_b200.engine import FockEngine
    eng = FockEngine.get(local_rank)
    nn, B = args.perm_n, args.batch
    mats_h = torch.from_numpy(haar_submatrices(nn, B)).pin_memory()
    mats = mats_h.to(eng.device)
    fp64 = measure_fp64(eng)
    out = {}

    def step(src=mats):
        out["perm"] = pdist.permanents_sharded(lambda ms_, g0, g1: eng.permanents(ms_, g0, g1), src)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    # parity of the sharded result: Gray-code prefix of the first matrix against the CPU oracle + all ranks agree
    got = out["perm"].cpu().numpy()
    if args.spot and rank == 0:
        import oracle
        G = 1 << (nn - 1)
        g1 = G >> 8
        part = complex(eng.permanents(mats[:1], 0, g1).cpu().numpy()[0])
        ref = oracle.permanent(mats_h[0].numpy(), 0, g1)
        assert abs(part - ref) <= 1e-10 * abs(ref), (part, ref)
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count()
    ms = timed_steps(torch, barrier, rmax, step, args.steps)
    launches = eng.launch_count() - launches0

    host_out = torch.empty(B, dtype=torch.complex128).pin_memory()

    def e2e_step():
        step(mats_h.to(eng.device, non_blocking=True))
        host_out.copy_(out["perm"], non_blocking=True)

    e2e_step()
    e2e_ms = timed_steps(torch, barrier, rmax, e2e_step, max(1, min(args.steps, 3)))
    clk = clocks.stop()
    assert np.allclose(host_out.numpy(), got, rtol=1e-12, atol=0)
    flops = B * perm_flops(nn)
    tf = flops / (ms * 1e-3) / 1e12
    roofline = {"kernel": f"glynn_big_kernel<{nn}> (csrc/permanent.cu)", "bound": "fp64", "achieved": tf, "peak": (fp64 or 0) * world,
                "peak_source": "FP64 FMA micro-benchmark measured in this run (fock_measure_peak, csrc/peaks.cu) x n_gpus; not in "
                               "MEASURED_PEAKS.json", "unit": "TFLOP/s", "frac": tf / (fp64 * world) if fp64 else None,
                "algorithmic_flops": flops, "traffic": None,
                "note": "8n-4 algorithmic flops per Gray step issue as 6n-4 FP64 instructions: 100 % FP64 pipe = 0.67 of the FMA peak"}
    line = {"metric": "glynn_permanents_per_s", "value": B / (ms * 1e-3), "unit": "permanents/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic", "config": dict(perm_config(nn, B), partition="by matrix (all-gather of one complex per matrix)" if B >= world
                                                else "by Gray-code range (all-reduce of one complex per matrix)"),
            "roofline": roofline,
            "e2e": {"value": B / (e2e_ms * 1e-3), "unit": "permanents/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(B * nn * nn * 16),
                    "d2h_bytes_per_step": int(B * 16), "api": "matrices from pinned host memory + dist.permanents_sharded + permanents to pinned host memory"},
            "gpu_launches": int(launches), "clocks": clk}
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_cc2017(args, torch, dist, world, rank, local_rank, barrier, rmax):
    from perceval_b200 import dist as pdist
    from perceval_b200.circuit import random_unitary
    from perceval_b200.engine import FockEngine
    eng = FockEngine.get(local_rank)
    m = args.modes if args.modes != N_MODES else 400
    n = args.photons if args.photons != N_PHOTONS else 20
    count = args.samples
    st = [1] * n + [0] * (m - n)
    u_np = random_unitary(m, seed=0)
    u_host = torch.from_numpy(u_np).pin_memory()
    U = eng.unitary(u_host)
    fp64 = measure_fp64(eng)
    b, e = pdist.shard_range(count, rank, world)
    buf = torch.empty((e - b, m), dtype=torch.uint8, device=eng.device)
    out = {}

    def step(Udev=U):
        out["smp"] = pdist.samples_sharded(lambda c, off: eng.cc2017_samples(Udev, st, c, seed=0, offset=off, out=buf[:c]), count, gather=False)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    if args.spot:   # the first samples of this rank's index range, bit for bit against the CPU oracle
        import oracle
        k = min(16, e - b)
        assert (out["smp"][:k].cpu().numpy() == oracle.cc2017_samples(u_np, tuple(st), k, seed=0, offset=b)).all()
    assert bool((out["smp"].sum(dim=1, dtype=torch.int32) == n).all().item())
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count()
    ms = timed_steps(torch, barrier, rmax, step, args.steps)
    launches = eng.launch_count() - launches0
    host = torch.empty((e - b, m), dtype=torch.uint8).pin_memory()

    def e2e_step():
        Ud = torch.empty_like(U)
        Ud.copy_(u_host, non_blocking=True)
        step(Ud)
        host.copy_(out["smp"], non_blocking=True)

    e2e_step()
    e2e_ms = timed_steps(torch, barrier, rmax, e2e_step, max(1, min(args.steps, 3)))
    clk = clocks.stop()
    flops = count * (sum(8.0 * k * 2.0 ** (k - 1) for k in range(1, n + 1)) + 8.0 * m * n * n)
    tf = flops / (ms * 1e-3) / 1e12
    roofline = {"kernel": "cc2017_kernel (csrc/cc2017.cu)", "bound": "fp64", "achieved": tf, "peak": (fp64 or 0) * world,
                "peak_source": "FP64 FMA micro-benchmark measured in this run (fock_measure_peak) x n_gpus; not in MEASURED_PEAKS.json",
                "unit": "TFLOP/s", "frac": tf / (fp64 * world) if fp64 else None, "algorithmic_flops": flops, "traffic": None,
                "note": "model: sum_k 8 k 2^(k-1) (Laplace sub-permanents, one Gray sweep per photon) + 8 m n^2 (weights) per sample"}
    line = {"metric": "cc2017_samples_per_s", "value": count / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128",
            "data": "synthetic", "config": dict(cc_config(n, m, count), partition="by sample index range (Philox stream keyed by the global index), samples stay on their GPU"),
            "roofline": roofline,
            "e2e": {"value": count / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(m * m * 16),
                    "d2h_bytes_per_step": int((e - b) * m), "api": "U from pinned host memory + dist.samples_sharded + this rank's samples to pinned host memory"},
            "gpu_launches": int(launches), "clocks": clk}
    if rank == 0:
        print(json.dumps(line), flush=True)


def extras(eng, torch):
    """Single-GPU figures for the other BASELINE configs (permanents n = 24 / 30 / 32, C&C sampling 20 photons / 400 modes)."""
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

    for nn, B, reps in [(24, 64, 3), (30, 8, 2), (32, 8, 1)]:
        mats = torch.from_numpy(haar_submatrices(nn, B)).to(eng.device)
        ms = timed(lambda: eng.permanents(mats), reps)
        flops = B * perm_flops(nn)
        rec = {"per_s": B / (ms * 1e-3), "batch": B, "ms": ms, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12}
        if fp64:
            rec["frac_of_measured_fp64_peak"] = rec["algorithmic_tflops"] / fp64
            # the pipe issues 6n-4 DFMA/DMUL per Gray step for 8n-4 algorithmic flops
            rec["fp64_pipe_utilisation_est"] = (B * (2.0 ** (nn - 1)) * (6 * nn - 4) * 2) / (ms * 1e-3) / 1e12 / fp64
        out[f"permanents_n{nn}"] = rec
    m, n, count = 400, 20, 100000
    U = eng.unitary(random_unitary(m, seed=0))
    st = [1] * n + [0] * (m - n)
    buf = torch.empty((count, m), dtype=torch.uint8, device=eng.device)
    ms = timed(lambda: eng.cc2017_samples(U, st, count, seed=0, out=buf), 1)
    flops = count * (sum(8.0 * k * 2.0 ** (k - 1) for k in range(1, n + 1)) + 8.0 * m * n * n)
    out["cc2017_n20_m400"] = {"samples_per_s": count / (ms * 1e-3), "count": count, "ms": ms, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                              "frac_of_measured_fp64_peak": flops / (ms * 1e-3) / 1e12 / fp64 if fp64 else None}
    return out


def run_b200(args):
    torch, dist, world, rank, local_rank, barrier, rmax = dist_setup()
    try:
        if args.workload == "permanents":
            run_permanents(args, torch, dist, world, rank, local_rank, barrier, rmax)
        elif args.workload == "cc2017":
            run_cc2017(args, torch, dist, world, rank, local_rank, barrier, rmax)
        elif world == 1:
            run_slos_single(args, torch, rank, local_rank, barrier, rmax)
        else:
            run_slos_multi(args, torch, dist, world, rank, local_rank, barrier, rmax)
    finally:
        if world > 1:
            dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="slos", choices=["slos", "permanents", "cc2017"])
    ap.add_argument("--photons", type=int, default=N_PHOTONS)
    ap.add_argument("--modes", type=int, default=N_MODES)
    ap.add_argument("--perm-n", type=int, default=30, help="permanents: matrix size (torchrun's own parser swallows a bare --n)")
    ap.add_argument("--batch", type=int, default=8, help="permanents: matrices per step (whole job)")
    ap.add_argument("--samples", type=int, default=100000, help="cc2017: samples per step (whole job)")
    ap.add_argument("--e2e-pieces", type=int, default=8)
    ap.add_argument("--partition", default="auto", choices=["auto", "replicate", "exchange", "slab", "windowed"],
                    help="N > 1: slab = prefix slabs owned by rank, prefix rows over NVLink (default when two whole layers fit); "
                         "replicate = lower layers on every rank, output layer sharded by rank range; exchange = rank ranges + NVLink "
                         "halo exchange; windowed = recompute-window chain (no exchange, the only one that fits 14/28)")
    ap.add_argument("--pieces", type=int, default=0, help="slab / exchange: pieces per rank and layer = exchange groups per layer (0: slab takes 4 with >= 10 prefix modes, where the windows of consecutive pieces are nearly disjoint -- 14/28 --, else 1; exchange takes 4)")
    ap.add_argument("--shard-min", type=int, default=1 << 23, help="exchange: layers with fewer states are replicated")
    ap.add_argument("--sub", type=int, default=0, help="windowed: sub-shards per rank (0 = from free memory)")
    ap.add_argument("--spot", type=int, default=64, help="outputs (per rank) checked against the CPU oracle outside the timed region")
    ap.add_argument("--reference-budget-s", type=float, default=240.0)
    ap.add_argument("--timeline", action="store_true", help="slab partition: per-rank time of every phase of one step (CUDA events)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
