"""Multi-GPU partitioning of the Fock-amplitude path: one process per GPU, torch.distributed (NCCL over NVLink) plumbing.

* SLOS -- owner-computes on the child layer.  Every rank owns a contiguous rank range of each child layer; the parent
  layer must be resident where it is read, so after each intermediate layer the shards are exchanged with ONE
  all-gather (``exchange="allgather"``, the scheme BASELINE.json's north_star names).  Because NVLink (~0.7 TB/s
  all-gather bus bandwidth) is slower than re-computing a layer from HBM, ``exchange="replicate"`` instead lets every
  rank recompute the (small) intermediate layers redundantly with no communication at all; ``"auto"`` picks per layer
  from a bandwidth model.  The last layer (two thirds of all traffic) is always sharded and stays sharded; only the
  scalar sum(p) is all-reduced.  Results are independent of the world size bit for bit (each child is computed by the
  same kernel from the same parent values).
* permanents -- a batch is split by matrix; a single large permanent is split by Gray-code range with one all-reduce.
* sampling -- the Philox stream is keyed by the global sample index, so ranks draw disjoint index ranges and an optional
  all-gather restores the chronological order.

The compute step is injected (``layer_fn`` ...) so that the partition / exchange logic is testable on CPU with gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of [0, total): the first (total % world) ranks get one extra element."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_ragged(shard: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather of balanced contiguous shards (see shard_range) of a 1-D tensor of length ``total``."""
    rank, world = _world(group)
    if world == 1:
        return shard
    is_complex = shard.is_complex()
    flat = torch.view_as_real(shard).reshape(-1) if is_complex else shard.reshape(-1)
    width = 2 if is_complex else 1
    maxlen = (total + world - 1) // world * width
    send = flat
    if flat.numel() < maxlen:
        send = torch.zeros(maxlen, dtype=flat.dtype, device=flat.device)
        send[:flat.numel()] = flat
    recv = torch.empty(world * maxlen, dtype=flat.dtype, device=flat.device)
    try:
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        chunks = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(chunks, send.contiguous(), group=group)
        recv = torch.cat(chunks)
    if total % world == 0:
        out = recv
    else:
        parts = []
        for r in range(world):
            b, e = shard_range(total, r, world)
            parts.append(recv[r * maxlen: r * maxlen + (e - b) * width])
        out = torch.cat(parts)
    return torch.view_as_complex(out.view(-1, 2)) if is_complex else out


# measured-ish bandwidth model (GB/s) used by exchange="auto"; see DESIGN.md section 6
HBM_EFFECTIVE_GBS = 2500.0
NVLINK_ALLGATHER_GBS = 700.0


def choose_exchange(n_parent: int, n_child: int, world: int) -> str:
    """'replicate' if recomputing the whole child layer on every rank is cheaper than all-gathering its shards."""
    if world == 1:
        return "replicate"
    extra_compute = 16.0 * (n_parent + n_child) * (1.0 - 1.0 / world) / HBM_EFFECTIVE_GBS
    gather = 16.0 * n_child * (world - 1) / world / NVLINK_ALLGATHER_GBS
    return "replicate" if extra_compute <= gather else "allgather"


def slos_probs_sharded(m: int, in_state, order, count_fn, layer_fn, last_fn, group=None, exchange: str = "auto"):
    """Layer-sharded SLOS chain.

    count_fn(m, k) -> N(k);  layer_fn(k, mk, parent_full, begin, end) -> child[begin:end] (complex128 1-D tensor);
    last_fn(k, mk, parent_full, begin, end) -> (probs[begin:end], partial_sum tensor of 1 element).
    Returns (probs_shard, (begin, end), total_sum tensor, exchanges) where exchanges lists the decision per layer."""
    rank, world = _world(group)
    n = sum(int(x) for x in in_state)
    decisions = []
    parent = None  # layer 0 is created by the first layer_fn call (parent_full=None means the vacuum coefficient [1])
    for k in range(1, n):
        nc, npar = count_fn(m, k), count_fn(m, k - 1)
        mode = exchange if exchange != "auto" else choose_exchange(npar, nc, world)
        if world == 1:
            mode = "replicate"
        decisions.append(mode)
        if mode == "replicate":
            parent = layer_fn(k, order[k - 1], parent, 0, nc)
        else:
            b, e = shard_range(nc, rank, world)
            shard = layer_fn(k, order[k - 1], parent, b, e)
            parent = all_gather_ragged(shard, nc, group)
    N = count_fn(m, n)
    b, e = shard_range(N, rank, world)
    probs, psum = last_fn(n, order[n - 1], parent, b, e)
    if world > 1:
        dist.all_reduce(psum, op=dist.ReduceOp.SUM, group=group)
    return probs, (b, e), psum, decisions


def engine_slos_probs_sharded(engine, U, in_state, group=None, exchange: str = "auto"):
    """The device instantiation of slos_probs_sharded (kernels from libfock_b200.so, NCCL all-gather)."""
    from .engine import prodnfact
    occ = [int(x) for x in in_state]
    m, n = len(occ), sum(occ)
    assert n >= 1
    order = engine.slos_order(occ)
    inf = prodnfact(occ)

    def vac():
        return torch.ones(1, dtype=torch.complex128, device=engine.device)

    def layer_fn(k, mk, parent, b, e):
        parent = vac() if parent is None else parent
        return engine.slos_layer(m, k, U, mk, parent, child_begin=b, child_end=e)

    def last_fn(k, mk, parent, b, e):
        parent = vac() if parent is None else parent
        psum = torch.zeros(1, dtype=torch.float64, device=engine.device)
        probs = engine.slos_layer_probs(m, k, U, mk, parent, inf, psum=psum, child_begin=b, child_end=e)
        return probs, psum

    return slos_probs_sharded(m, occ, order, engine.count, layer_fn, last_fn, group, exchange)


def windowed_plan(m: int, n: int, rank: int, world: int, sub: int = 1, balanced: bool = True, last_weight: float | None = None):
    """Sub-shards of a rank in the recompute-window partition: the output layer is cut in world * sub contiguous ranges
    of about equal chain cost (partition.balanced_boundaries; equal counts if ``balanced`` is False), rank r takes the
    ranges r*sub .. r*sub+sub-1, each with its own chain plan.  Returns [(begin, end, plan), ...]."""
    from . import partition as P
    N = P.count(m, n)
    total = world * sub
    lw = P.LAST_LAYER_WEIGHT if last_weight is None else last_weight
    bounds = P.balanced_boundaries(m, n, total, last_weight=lw) if balanced else [shard_range(N, i, total)[0] for i in range(total)] + [N]
    out = []
    for i in range(rank * sub, rank * sub + sub):
        b, e = bounds[i], bounds[i + 1]
        out.append((b, e, P.plan_chain(m, n, b, e)))
    return out


def windowed_buffer_elems(n: int, pieces):
    """(A, B): complex128 elements of the two packed ping-pong layer buffers that serve every sub-shard of ``pieces``
    (layers n-1, n-3, ... live in A, layers n-2, n-4, ... in B)."""
    from . import partition as P
    a = max([P.segments_len(plan[k]) for _, _, plan in pieces for k in range(n - 1, -1, -2)] + [1])
    b = max([P.segments_len(plan[k]) for _, _, plan in pieces for k in range(n - 2, -1, -2)] + [1])
    return a, b


def tail_table_bytes(n: int, D: int = 16) -> int:
    """Device bytes of the cached tail occupation tables (csrc/slos_mu.cu): 8 B per tail rank of FS(D, u), u <= n."""
    from . import partition as P
    return 8 * P.count(D + 1, n)


def windowed_peak_bytes(n: int, pieces) -> int:
    """Device bytes of a WindowedChain: both layer buffers, the probabilities of the rank's whole range and the library's
    cached tail tables (1.2 GB at 14 photons)."""
    a, b = windowed_buffer_elems(n, pieces)
    return 16 * (a + b) + 8 * sum(e - bb for bb, e, _ in pieces) + tail_table_bytes(n)


def windowed_candidates(sub: int | None = None):
    """(sub-shards per rank, balance weight) in order of expected speed: fewer sub-shards recompute less; the time-weighted
    balance (output-layer states cost ~3x) is faster but gives the middle ranks wider windows than the state-count one."""
    from . import partition as P
    subs = [sub] if sub is not None else list(range(1, 65))
    return [(s_, w) for s_ in subs for w in (P.LAST_LAYER_WEIGHT, 1.0)]


def windowed_pick(m: int, n: int, rank: int, world: int, free_bytes: int, mem_fraction: float = 0.85, sub: int | None = None,
                  group=None, collective: bool = True) -> tuple:
    """First candidate whose workspace fits ``mem_fraction * free_bytes`` on this rank, then the LAST such index over all
    ranks (all-reduce MAX) so that every rank cuts the output layer at the same boundaries."""
    cands = windowed_candidates(sub)
    pick = len(cands) - 1
    for i, (s_, w) in enumerate(cands):
        if windowed_peak_bytes(n, windowed_plan(m, n, rank, world, s_, last_weight=w)) <= mem_fraction * free_bytes:
            pick = i
            break
    if collective and world > 1 and dist.is_available() and dist.is_initialized():
        t = torch.tensor([pick], dtype=torch.int64)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        pick = int(t.item())
    return cands[pick]


class WindowedChain:
    """Recompute-window partition of the SLOS chain: NO exchange step.  Every rank owns a contiguous range of the output
    layer and recomputes, layer by layer, exactly the parents that range needs (one or two rank ranges per layer,
    partition.py), so no layer is ever replicated or gathered.  This is what lets 14 photons / 28 modes (layer 13 alone is
    192 GB) run on 8 x 180 GB.  The rank's range is cut in ``sub`` sub-shards processed one after the other to bound the
    layer buffers (None: the smallest count whose workspace fits ``mem_fraction`` of the free device memory).  Buffers
    and plans are built once; ``run(U)`` is one step."""

    def __init__(self, engine, in_state, group=None, sub: int | None = None, mem_fraction: float = 0.85, as_rank=None):
        self.engine = engine
        self.occ = [int(x) for x in in_state]
        self.m, self.n = len(self.occ), sum(self.occ)
        assert self.n >= 1
        self.group = group
        self.rank, self.world = _world(group)
        emulated = as_rank is not None   # (rank, world) of a run emulated on this device alone (tools, tests): no collectives
        if emulated:
            self.rank, self.world = as_rank
        free, _total = torch.cuda.mem_get_info(engine.device)
        sub, self.last_weight = windowed_pick(self.m, self.n, self.rank, self.world, free, mem_fraction, sub, group,
                                              collective=not emulated)
        self.sub = sub
        self.pieces = windowed_plan(self.m, self.n, self.rank, self.world, sub, last_weight=self.last_weight)
        self.begin, self.end = self.pieces[0][0], self.pieces[-1][1]
        a, b = windowed_buffer_elems(self.n, self.pieces)
        self.bytes = windowed_peak_bytes(self.n, self.pieces)
        self.buffers = (torch.empty(a, dtype=torch.complex128, device=engine.device),
                        torch.empty(b, dtype=torch.complex128, device=engine.device))
        self.probs = torch.empty(self.end - self.begin, dtype=torch.float64, device=engine.device)
        self.psum = torch.zeros(1, dtype=torch.float64, device=engine.device)
        self.emulated = emulated

    def run(self, U, reduce_sum: bool = True, last_events: list | None = None):
        """one step: returns (probabilities of [begin, end), (begin, end), sum(p) over all ranks)"""
        self.psum.zero_()
        for b, e, plan in self.pieces:
            self.engine.slos_probs_windowed(U, self.occ, b, e, probs=self.probs[b - self.begin:e - self.begin], psum=self.psum,
                                            plan=plan, buffers=self.buffers, last_events=last_events)
        if reduce_sum and self.world > 1 and not self.emulated:
            dist.all_reduce(self.psum, op=dist.ReduceOp.SUM, group=self.group)
        return self.probs, (self.begin, self.end), self.psum


def engine_slos_probs_windowed(engine, U, in_state, group=None, sub: int | None = None):
    """One-shot form of WindowedChain."""
    return WindowedChain(engine, in_state, group, sub).run(U)


def _intersect(segs, b: int, e: int):
    """parts of the half-open ranges ``segs`` inside [b, e), merged"""
    out = []
    for lo, hi in segs:
        lo, hi = max(lo, b), min(hi, e)
        if hi > lo:
            if out and lo <= out[-1][1]:
                out[-1] = (out[-1][0], max(out[-1][1], hi))
            else:
                out.append((lo, hi))
    return out


def split_range(b: int, e: int, pieces: int):
    pieces = max(1, min(pieces, e - b)) if e > b else 1
    return [(b + (e - b) * i // pieces, b + (e - b) * (i + 1) // pieces) for i in range(pieces)]


class ExchangePlan:
    """Host-side plan of the owner-computes + halo-exchange partition of a SLOS chain.

    Layers below ``k0`` are replicated (they are small: recomputing them costs less than one message); every layer
    k >= k0 is cut in ``world`` contiguous rank ranges, rank q OWNS range q of every such layer and computes only that.
    The parents of layer k-1 that the children [b, e) need are one or two contiguous rank ranges (partition.parent_segments,
    exact), so what rank r must receive from rank q is a short list of slices of q's own range -- no whole-layer all-gather,
    no recompute.  The own range of a rank is cut in ``pieces`` child pieces, computed in order; the exchange that follows a
    layer is cut in the same number of GROUPS, group g carrying exactly the parents that child piece g of the next layer
    needs and no earlier group carried.  Child piece g therefore starts as soon as group g has arrived, while groups g+1..
    are still on the wire: the NVLink transfers of a layer overlap the computation of the next one.  Everything is integer
    arithmetic, identical on every rank.
    """

    def __init__(self, m: int, n: int, world: int, pieces: int = 4, shard_min: int = 1 << 23, max_segments: int = 4,
                 fractions=None):
        from . import partition as P
        self.m, self.n, self.world, self.pieces = m, n, world, pieces
        cnt = [P.count(m, k) for k in range(n + 1)]
        self.count = cnt
        k0 = n
        for k in range(1, n + 1):
            if cnt[k] >= shard_min:
                k0 = k
                break
        self.k0 = max(1, min(k0, n))
        self.fractions = fractions      # optional rank-space cut points [0, f1, .., 1]; None: equal counts
        self.own = {}      # k -> [(b, e)] per rank
        self.piece = {}    # k -> per rank, [(b, e)] per piece
        for k in range(self.k0, n + 1):
            if fractions is not None:
                bd = [0] + [min(cnt[k], max(0, int(round(f * cnt[k])))) for f in fractions[1:-1]] + [cnt[k]]
                for q in range(1, len(bd)):
                    bd[q] = max(bd[q], bd[q - 1])
                self.own[k] = [(bd[q], bd[q + 1]) for q in range(world)]
            else:
                self.own[k] = [shard_range(cnt[k], q, world) for q in range(world)]
            self.piece[k] = [split_range(b, e, pieces) for b, e in self.own[k]]
        # need[k][q][j]: parent (layer k-1) segments of child piece j of rank q; fresh[k][q][j]: the part no earlier piece needs
        self.need, self.fresh = {}, {}
        for k in range(self.k0 + 1, n + 1):
            self.need[k] = [[P.parent_segments(m, k, [pc], max_segments) if pc[1] > pc[0] else [] for pc in self.piece[k][q]]
                            for q in range(world)]
            self.fresh[k] = []
            for q in range(world):
                seen, out = [], []
                for segs in self.need[k][q]:
                    out.append(_subtract(segs, seen))
                    seen = P.merge_segments(seen + list(segs), max_segments=1 << 30)
                self.fresh[k].append(out)

    def npieces(self, k: int, q: int) -> int:
        return len(self.piece[k][q])

    def groups(self, k: int) -> int:
        """number of exchange groups after layer k (k0 <= k < n) = the largest child piece count of layer k+1"""
        return max(len(p) for p in self.piece[k + 1])

    def transfers(self, k: int, g: int, src: int, dst: int):
        """slices of layer k (k0 <= k < n) that rank ``src`` sends to ``dst`` in group g: the part of src's own range that
        child piece g of dst at layer k+1 needs and no earlier piece of dst needed"""
        if src == dst or g >= len(self.fresh[k + 1][dst]):
            return []
        b, e = self.own[k][src]
        return _intersect(self.fresh[k + 1][dst][g], b, e)

    def recv_elems(self, r: int, k: int | None = None) -> int:
        """complex elements rank r receives per step (after layer k only, if given)"""
        tot = 0
        for kk in ([k] if k is not None else range(self.k0, self.n)):
            for g in range(self.groups(kk)):
                for q in range(self.world):
                    tot += sum(hi - lo for lo, hi in self.transfers(kk, g, q, r))
        return tot

    def send_elems(self, r: int, k: int | None = None) -> int:
        tot = 0
        for kk in ([k] if k is not None else range(self.k0, self.n)):
            for g in range(self.groups(kk)):
                for q in range(self.world):
                    tot += sum(hi - lo for lo, hi in self.transfers(kk, g, r, q))
        return tot


def _subtract(segs, seen):
    """parts of ``segs`` not covered by the sorted, merged ranges ``seen``"""
    out = []
    for lo, hi in segs:
        cur = lo
        for a, b in seen:
            if b <= cur:
                continue
            if a >= hi:
                break
            if a > cur:
                out.append((cur, a))
            cur = max(cur, b)
            if cur >= hi:
                break
        if cur < hi:
            out.append((cur, hi))
    return out


class ExchangeChain:
    """One SLOS chain under ExchangePlan on this rank.  The compute step is injected so that the exchange logic runs on CPU
    with gloo in the tests:

      layer_fn(k, mk, parent_full, out, b, e)        writes child ranks [b, e) of layer k into ``out`` (length e - b);
                                                      parent_full is indexed by absolute rank (None: the vacuum, k = 1)
      last_fn(k, mk, parent_full, out, psum, b, e)   same for the output layer: float64 probabilities + partial sum

    Layers live in two full-size ping-pong buffers indexed by absolute rank (at 12 photons / 24 modes: 4.6 + 1.5 GB), so the
    kernels run their whole-parent fast paths; only the ranks a rank owns or receives are ever written or read.
    """

    def __init__(self, m: int, in_state, order, plan: ExchangePlan, alloc, layer_fn, last_fn, group=None, alloc_real=None):
        self.m = m
        self.occ = [int(x) for x in in_state]
        self.n = sum(self.occ)
        assert self.n >= 1 and plan.n == self.n and plan.m == m
        self.order = order
        self.plan = plan
        self.group = group
        self.rank, self.world = _world(group)
        assert self.world == plan.world, "plan built for another world size"
        self.layer_fn, self.last_fn = layer_fn, last_fn
        n = self.n
        cnt = plan.count
        self.buf_a = alloc(max(cnt[n - 1], 1))                       # layers n-1, n-3, ...
        self.buf_b = alloc(max(cnt[n - 2], 1) if n >= 2 else 1)      # layers n-2, n-4, ...
        b, e = plan.own[n][self.rank]
        self.begin, self.end = b, e
        self.probs = (alloc_real or (lambda k: torch.empty(k, dtype=torch.float64, device=self.buf_a.device)))(max(e - b, 1))[:e - b]
        self.psum = torch.zeros(1, dtype=torch.float64, device=self.buf_a.device)
        # static per-step schedule: for every exchange group the (peer, slice) lists of this rank
        r = self.rank
        self._xfer = {}
        for k in range(plan.k0, n):
            for g in range(plan.groups(k)):
                sends = [(q, seg) for q in range(self.world) for seg in plan.transfers(k, g, r, q)]
                recvs = [(q, seg) for q in range(self.world) for seg in plan.transfers(k, g, q, r)]
                self._xfer[(k, g)] = (sends, recvs)
        self.bytes_received = 16 * plan.recv_elems(r)
        self.bytes_sent = 16 * plan.send_elems(r)

    def _buf(self, k: int):
        return self.buf_a if (self.n - 1 - k) % 2 == 0 else self.buf_b

    def _exchange(self, k: int, g: int, buf):
        sends, recvs = self._xfer[(k, g)]
        if not sends and not recvs:
            return None
        flat = torch.view_as_real(buf)
        ops = []
        for q, (lo, hi) in recvs:
            ops.append(dist.P2POp(dist.irecv, flat[lo:hi], q, self.group))
        for q, (lo, hi) in sends:
            ops.append(dist.P2POp(dist.isend, flat[lo:hi], q, self.group))
        return [dist.batch_isend_irecv(ops), False]

    @staticmethod
    def _wait(group):
        """wait once for an exchange group (a second wait() on a completed gloo send / recv never returns)"""
        if group is None or group[1]:
            return
        for w in group[0]:
            w.wait()
        group[1] = True

    def run(self, reduce_sum: bool = True, on_last_piece=None):
        """one step: returns (probabilities of [begin, end), (begin, end), sum(p))"""
        plan, n, r = self.plan, self.n, self.rank
        cnt = plan.count
        self.psum.zero_()
        parent = None
        for k in range(1, plan.k0):                                   # replicated layers
            buf = self._buf(k)
            self.layer_fn(k, self.order[k - 1], parent, buf[:cnt[k]], 0, cnt[k])
            parent = buf
        groups = []                                                   # exchange groups that followed layer k-1, in order
        older = []                                                    # ... and layer k-2: its buffer is about to be rewritten
        for k in range(plan.k0, n + 1):
            buf = self._buf(k) if k < n else None
            for w in older:                                           # (always complete by now; keeps the re-use explicit)
                self._wait(w)
            for j, (b, e) in enumerate(plan.piece[k][r]):
                if j < len(groups):
                    self._wait(groups[j])                             # the parents child piece j needs have arrived
                if e <= b:
                    continue
                if k < n:
                    self.layer_fn(k, self.order[k - 1], parent, buf[b:e], b, e)
                else:
                    if on_last_piece is not None:
                        on_last_piece(j, "begin")
                    self.last_fn(k, self.order[k - 1], parent, self.probs[b - self.begin:e - self.begin], self.psum, b, e)
                    if on_last_piece is not None:
                        on_last_piece(j, "end")
            for w in groups[len(plan.piece[k][r]):]:
                self._wait(w)
            older = groups
            groups = [self._exchange(k, g, buf) for g in range(plan.groups(k))] if k < n else []
            parent = buf
        for w in older:
            self._wait(w)
        if reduce_sum and self.world > 1:
            dist.all_reduce(self.psum, op=dist.ReduceOp.SUM, group=self.group)
        return self.probs, (self.begin, self.end), self.psum


def engine_exchange_chain(engine, U_ref, in_state, group=None, pieces: int = 4, shard_min: int = 1 << 23, fractions=None):
    """Device instantiation of ExchangeChain: kernels from libfock_b200.so, NCCL send / recv over NVLink.  ``U_ref`` is a
    one-element list holding the device unitary, so that a step can swap it without rebuilding plan and buffers."""
    from .engine import prodnfact
    occ = [int(x) for x in in_state]
    m, n = len(occ), sum(occ)
    _rank, world = _world(group)
    order = engine.slos_order(occ)
    inf = prodnfact(occ)
    plan = ExchangePlan(m, n, world, pieces=pieces, shard_min=shard_min, fractions=fractions)
    vac = torch.ones(1, dtype=torch.complex128, device=engine.device)

    def alloc(k):
        return torch.empty(k, dtype=torch.complex128, device=engine.device)

    def layer_fn(k, mk, parent, out, b, e):
        engine.slos_layer(m, k, U_ref[0], mk, vac if parent is None else parent[:plan.count[k - 1]], child=out, child_begin=b, child_end=e)

    def last_fn(k, mk, parent, out, psum, b, e):
        engine.slos_layer_probs(m, k, U_ref[0], mk, vac if parent is None else parent[:plan.count[k - 1]], inf, probs=out, psum=psum,
                                child_begin=b, child_end=e)

    return ExchangeChain(m, occ, order, plan, alloc, layer_fn, last_fn, group)


def permanents_sharded(perm_fn, mats: torch.Tensor, group=None):
    """Batch of permanents over ranks.  perm_fn(mats, gray_begin, gray_end) -> (B,) complex tensor (partial sums for a
    Gray sub-range, already scaled).  B >= world: split by matrix + all-gather; else split the Gray range + all-reduce."""
    rank, world = _world(group)
    B, n = mats.shape[0], mats.shape[1]
    if world == 1:
        return perm_fn(mats, 0, 0)
    if B >= world:
        b, e = shard_range(B, rank, world)
        part = perm_fn(mats[b:e], 0, 0)
        return all_gather_ragged(part, B, group)
    G = 1 << max(n - 1, 0)
    b, e = shard_range(G, rank, world)
    part = perm_fn(mats, b, e) if e > b else torch.zeros(B, dtype=torch.complex128, device=mats.device)
    buf = torch.view_as_real(part).contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return torch.view_as_complex(buf)


def samples_sharded(sample_fn, count: int, offset: int = 0, group=None, gather: bool = True):
    """sample_fn(count, offset) -> (count, m) uint8 tensor of the samples with global indices [offset, offset+count)."""
    rank, world = _world(group)
    b, e = shard_range(count, rank, world)
    part = sample_fn(e - b, offset + b)
    if world == 1 or not gather:
        return part
    m = part.shape[1]
    flat = all_gather_ragged(part.reshape(-1), count * m, group) if (count % world == 0) else None
    if flat is None:
        # ragged in units of whole samples: gather per-sample padded
        maxc = (count + world - 1) // world
        send = torch.zeros((maxc, m), dtype=part.dtype, device=part.device)
        send[: e - b] = part
        chunks = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(chunks, send, group=group)
        outs = []
        for r in range(world):
            rb, re_ = shard_range(count, r, world)
            outs.append(chunks[r][: re_ - rb])
        return torch.cat(outs)
    return flat.view(count, m)
