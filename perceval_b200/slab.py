"""Slab partition of the SLOS chain for several GPUs: owner-computes on prefix slabs, halo = parent rows only.

Layout.  The m modes are split, exactly as in the tile kernels, into a PREFIX (first p = m - D modes) and a TAIL (last D).
A layer k is stored SLAB-MAJOR: slab w (= photons in the prefix, 0..k) holds, for every prefix rho of FS(p, w) in FSArray
order, the block of the S(k, w) = |FS(D, k - w)| tails in FSArray order:

    index(k; w, rho, t) = slab_off[k][w] + rho * S(k, w) + t

(a permutation of the FSArray rank order by whole prefix blocks; `rank_to_slab` / `slab_to_rank` convert).  One SLOS step
then reads, for child (w, rho, t):  tail-mode parents from block (w, rho) of layer k-1 -- the SAME prefix -- and
prefix-mode parents from the aligned rows (w-1, rho', t), rho' = rank of (prefix - e_j) in FS(p, w-1)
(reference perceval/backends/_slos.py:91-97 read as a gather; kernels: slos_layer_slab in csrc/slos.cu).

Ownership.  Rank q owns a contiguous run of prefixes in slab-major order (w ascending, rho ascending), THE SAME for every
layer, cut so that the states of all sharded layers together are balanced.  Tail-mode parents are therefore always local; what crosses NVLink are
prefix-mode rows only: for every owned (w, [a, b)) the rho' window of slab w-1 (partition.parent_segments on the prefix
space FS(p, w), exact) minus what the rank owns itself -- contiguous slices of the neighbours' slabs.  At 12 photons /
24 modes on 8 GPUs a rank receives 1.0 GB for the output layer instead of the 1.6 - 3.0 GB of rank-contiguous ranges, every
rank sends about what it receives (no hot sender), and nothing is recomputed.
"""
from __future__ import annotations

from . import partition as P


def tail_modes(m: int) -> int:
    """tail width of the tile kernels (slos_tail_modes in csrc/slos.cu); 0: no tile kernel for this m"""
    D = 16 if m >= 20 else (12 if m >= 16 else (8 if m >= 12 else (6 if m >= 8 else (4 if m >= 6 else 0))))
    return D if D <= m - 1 else 0


class SlabLayout:
    def __init__(self, m: int, n: int):
        self.m, self.n = m, n
        self.D = tail_modes(m)
        assert self.D > 0, f"slab layout needs m >= 6 (m = {m})"
        self.p = m - self.D
        self.nprefix = [P.count(self.p, w) for w in range(n + 1)]
        # S[k][w], slab_off[k][w] for k = 0..n
        self.S = [[P.count(self.D, k - w) for w in range(k + 1)] for k in range(n + 1)]
        self.off = []
        for k in range(n + 1):
            o, acc = [], 0
            for w in range(k + 1):
                o.append(acc)
                acc += self.nprefix[w] * self.S[k][w]
            assert acc == P.count(m, k)
            self.off.append(o)

    def index(self, k: int, state) -> int:
        """slab-major index of an occupation tuple of layer k"""
        pre, tail = list(state[:self.p]), list(state[self.p:])
        w = sum(pre)
        return self.off[k][w] + P.rank(self.p, pre) * self.S[k][w] + P.rank(self.D, tail)

    def rank_to_slab(self, k: int, ranks):
        """numpy int64 slab-major indices of FSArray ranks of layer k (vectorised over whole prefix blocks)"""
        import numpy as np
        ranks = np.asarray(ranks, dtype=np.int64)
        perm = self.permutation(k)
        return perm[ranks]

    def permutation(self, k: int):
        """perm[rank] = slab-major index, for the whole layer k (host memory: 8 B per state -- tests and spot checks only)"""
        import numpy as np
        key = ("perm", k)
        if not hasattr(self, "_cache"):
            self._cache = {}
        if key not in self._cache:
            N = P.count(self.m, k)
            perm = np.empty(N, dtype=np.int64)
            r = 0
            # FSArray order = prefixes in descending lexicographic order over ALL weights, each followed by its tail block
            for pre in _prefixes_desc(self.p, k):
                w = sum(pre)
                S = self.S[k][w]
                base = self.off[k][w] + P.rank(self.p, pre) * S
                perm[r:r + S] = np.arange(base, base + S, dtype=np.int64)
                r += S
            assert r == N
            self._cache[key] = perm
        return self._cache[key]


def _subtract(segs, seen):
    """parts of the half-open ranges ``segs`` not covered by the sorted, merged ranges ``seen``"""
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


def _prefixes_desc(p: int, kmax: int):
    """all prefixes of p modes with at most kmax photons, in the order they appear in FSArray(m, k) (descending lex)"""
    def rec(i, left):
        if i == p:
            yield []
            return
        for a in range(left, -1, -1):
            for rest in rec(i + 1, left - a):
                yield [a] + rest
    yield from rec(0, kmax)


class SlabPlan:
    """Ownership + halo lists, identical on every rank (pure integer arithmetic)."""

    def __init__(self, m: int, n: int, world: int, shard_min: int = 1 << 23, pieces: int = 4):
        self.layout = L = SlabLayout(m, n)
        self.m, self.n, self.world = m, n, world
        cnt = [P.count(m, k) for k in range(n + 1)]
        self.count = cnt
        k0 = n
        for k in range(1, n + 1):
            if cnt[k] >= shard_min:
                k0 = k
                break
        self.k0 = max(1, min(k0, n))
        # contiguous runs of whole prefixes, own[q] = [(w, a, b)], balanced on the modelled time a prefix costs over ALL the sharded
        # layers.  A child costs c + s / (prefixes its CTA sweeps): every thread sets its tail up once per sweep, and the slabs
        # with few prefixes (w = 0: one prefix, 2 % of the states of 12/24) amortise that over little work.  c = 12.2 ps,
        # s = 61 ps fitted to the measured single-GPU layer (14.6 ps per state) and to a rank that held only w <= 2 (21 ps).
        def per_state(w):
            if w == 0 and L.S[n][0] >= (1 << 18) and tail_modes(L.D) > 0:
                return 22.0     # a whole weight-0 slab runs as a sub-layer on the tail modes (csrc/slos.cu), not as a one-prefix sweep
            return 12.2 + 61.0 / min(max(L.nprefix[w], 1), 128)
        cost = [int(1000 * per_state(w)) * sum(L.S[k][w] for k in range(max(self.k0, w), n + 1)) for w in range(n + 1)]
        total = sum(L.nprefix[w] * cost[w] for w in range(n + 1))

        def cut(targets):
            """contiguous runs of whole prefixes whose cumulated cost follows ``targets`` (cumulative, per rank)"""
            own = [[] for _ in range(world)]
            cum, q = 0, 0
            for w in range(n + 1):
                sw, a = cost[w], 0
                while a < L.nprefix[w]:
                    room = targets[q] - cum
                    take = min(L.nprefix[w] - a, max(1, -(-room // sw))) if (room > 0 or q == world - 1) else 0
                    if take == 0:
                        q += 1
                        continue
                    if own[q] and own[q][-1][0] == w and own[q][-1][2] == a:
                        own[q][-1] = (w, own[q][-1][1], a + take)
                    else:
                        own[q].append((w, a, a + take))
                    cum += take * sw
                    a += take
                    if cum >= targets[q] and q < world - 1:
                        q += 1
            return own

        def owners_of(own):
            owners = {}
            for q in range(world):
                for w, a, b in own[q]:
                    owners.setdefault(w, []).append((q, a, b))
            return owners

        def remote_rows(own, owners, q):
            """[(src, w-1, lo, hi)] prefix rows rank q reads and does not own"""
            rows = []
            for w, a, b in own[q]:
                if w >= 1:
                    for lo, hi in P.parent_segments(L.p, w, [(a, b)], max_segments=4):
                        for src, sa, sb in owners.get(w - 1, []):
                            if src != q:
                                x, y = max(lo, sa), min(hi, sb)
                                if y > x:
                                    rows.append((src, w - 1, x, y))
            return rows

        # a rank that waits for rows does less in the same time: the cost a rank is given = its share of (compute + everybody's
        # transfers) minus its own transfer time (NCCL send / recv between B200s: ~0.55 B / ps while a rank sends and receives)
        self.own = cut([total * (q + 1) // world for q in range(world)])
        for _ in range(3 if world > 1 else 0):
            owners = owners_of(self.own)
            comm = []
            for q in range(world):
                elems = sum((hi - lo) * sum(L.S[k][wp] for k in range(max(self.k0, wp), n)) for src, wp, lo, hi in remote_rows(self.own, owners, q))
                # with several pieces per layer the transfer hides behind the computation (14/28: 187 ms with the plain cost
                # balance, 201 ms with a quarter of the transfer charged, 221 ms with all of it): charge it only when un-pipelined
                comm.append(int(1000 * 16 * elems / 0.55) if pieces <= 1 else 0)
            share = (total + sum(comm)) // world
            acc, targets = 0, []
            for q in range(world):
                acc += max(share - comm[q], total // (8 * world))
                targets.append(acc)
            self.own = cut([t * total // max(targets[-1], 1) for t in targets])
        self._storage = {}
        self._owners = owners_of(self.own)
        # rows[q][w] = merged rho' segments of slab w-1 that q's part of slab w reads through prefix modes (owned or not)
        self.rows = [{} for _ in range(world)]
        for q in range(world):
            for w, a, b in self.own[q]:
                if w >= 1:
                    self.rows[q][w] = P.parent_segments(L.p, w, [(a, b)], max_segments=4)
        # PIECES.  A rank's prefixes are cut into about ``pieces`` runs of equal modelled cost, computed one launch after the
        # other; the exchange that follows a layer is cut into as many GROUPS, group g carrying the rows piece g of the next
        # layer reads that no earlier piece needed and the rank does not own.  Piece g starts when group g has arrived, while the
        # later groups are still on the wire.  Pieces whose rows are all local come first (they never wait).  With 12 prefix
        # modes (14 photons / 28 modes) the windows of consecutive pieces are almost disjoint, so the transfer of a layer
        # hides behind its own computation; with 8 (12/24) the first piece needs 60 - 80 % of the halo and little is hidden.
        self.piece = [[] for _ in range(world)]         # piece[q][g] = [(w, a, b)]
        self.fresh = [[] for _ in range(world)]         # fresh[q][g] = [(src, w-1, lo, hi)] rows to receive before piece g
        for q in range(world):
            # runs whose rows are all local form their own pieces (they never wait), the others are cut by modelled cost
            free_runs = [(w, a, b) for w, a, b in self.own[q] if not remote_rows([[(w, a, b)] if r == q else [] for r in range(world)], self._owners, q)]
            rest_runs = [x for x in self.own[q] if x not in free_runs]
            mine = sum((b - a) * cost[w] for w, a, b in self.own[q])
            out = []
            for runs in (free_runs, rest_runs):
                part = sum((b - a) * cost[w] for w, a, b in runs)
                if part == 0:
                    continue
                npc = max(1, round(pieces * part / max(mine, 1)))
                target = max(1, -(-part // npc))
                cur, acc = [], 0
                for w, a, b in runs:
                    while a < b:
                        room = target - acc
                        take = min(b - a, max(1, room // max(cost[w], 1)))
                        if cur and cur[-1][0] == w and cur[-1][2] == a:
                            cur[-1] = (w, cur[-1][1], a + take)
                        else:
                            cur.append((w, a, a + take))
                        acc += take * cost[w]
                        a += take
                        if acc >= target:
                            out.append(cur)
                            cur, acc = [], 0
                if cur:
                    out.append(cur)
            need = []
            for pc in out:
                rows = []
                for w, a, b in pc:
                    if w >= 1:
                        for lo, hi in P.parent_segments(L.p, w, [(a, b)], max_segments=4):
                            for src, sa, sb in self._owners.get(w - 1, []):
                                if src != q:
                                    x, y = max(lo, sa), min(hi, sb)
                                    if y > x:
                                        rows.append((src, w - 1, x, y))
                need.append(rows)
            # processing order: greedily the piece whose rows add the least to what has already been received (pieces with
            # local rows only come first; windows that nest -- the late prefixes of a slab read a subset of what the early ones
            # read -- are walked from the inside out, so every group carries about the same volume)
            def volume(rows, seen):
                return sum((y - x) * L.S[n - 1][wp] if wp <= n - 1 else 0
                           for src, wp, lo, hi in rows for x, y in _subtract([(lo, hi)], seen.get((src, wp), [])))
            seen, left = {}, list(range(len(out)))
            while left:
                i = min(left, key=lambda j: (volume(need[j], seen), j))
                left.remove(i)
                fr = []
                for src, wp, lo, hi in need[i]:
                    for x, y in _subtract([(lo, hi)], seen.get((src, wp), [])):
                        fr.append((src, wp, x, y))
                    seen[(src, wp)] = P.merge_segments(seen.get((src, wp), []) + [(lo, hi)], max_segments=1 << 30)
                self.piece[q].append(out[i])
                self.fresh[q].append(fr)

    def groups(self) -> int:
        return max(len(p_) for p_ in self.piece)

    def owner_ranges(self, w: int):
        return self._owners.get(w, [])

    def own_elems(self, k: int, q: int) -> int:
        return sum((b - a) * self.layout.S[k][w] for w, a, b in self.own[q] if w <= k)

    def storage(self, k: int, q: int):
        """Compact per-rank layout of layer k (k0-1 <= k <= n-1) on rank q: per prefix weight w the contiguous prefix range
        [lo, hi) the rank holds -- its own prefixes (tail parents of its children of layer k+1, and what it computes itself)
        together with the rows of slab w its children of slab w+1 read -- and the element offset ``base`` of (w, lo).
        Layer k0-1 (replicated, then permuted to slab-major) is held whole.  Returns ([(lo, hi, base)] for w = 0..k, size)."""
        key = (k, q)
        if key not in self._storage:
            L = self.layout
            out, acc = [], 0
            for w in range(k + 1):
                if k < self.k0:
                    lo, hi = 0, L.nprefix[w]
                else:
                    spans = [(a, b) for ww, a, b in self.own[q] if ww == w]
                    if k < self.n:
                        spans += self.rows[q].get(w + 1, [])
                    lo = min((x for x, _ in spans), default=0)
                    hi = max((y for _, y in spans), default=0)
                out.append((lo, hi, acc))
                acc += (hi - lo) * L.S[k][w]
            self._storage[key] = (out, acc)
        return self._storage[key]

    def offsets(self, k: int, q: int):
        """element offsets off[w] of rank q's compact layer k with index(w, rho, t) = off[w] + rho * S + t (may be negative)"""
        L = self.layout
        return [base - lo * L.S[k][w] for w, (lo, hi, base) in enumerate(self.storage(k, q)[0])]

    def transfers(self, k: int, src: int, dst: int, g: int | None = None):
        """[(src_lo, dst_lo, length)] element slices of layer k (k0 <= k < n) that ``src`` sends to ``dst`` after computing it,
        in the compact coordinates of either rank: the rows dst's children of layer k+1 read through prefix modes (of dst's
        piece g only, if given)"""
        L = self.layout
        out = []
        so, do = self.offsets(k, src), self.offsets(k, dst)
        for gg, fr in enumerate(self.fresh[dst]):
            if g is not None and gg != g:
                continue
            for s_, wp, lo, hi in fr:
                if s_ == src and wp <= k:
                    S = L.S[k][wp]
                    out.append((so[wp] + lo * S, do[wp] + lo * S, (hi - lo) * S))
        return out

    def recv_elems(self, q: int, k: int | None = None) -> int:
        ks = [k] if k is not None else range(self.k0, self.n)
        return sum(ln for kk in ks for s_ in range(self.world) for _, _, ln in self.transfers(kk, s_, q))

    def send_elems(self, q: int, k: int | None = None) -> int:
        ks = [k] if k is not None else range(self.k0, self.n)
        return sum(ln for kk in ks for d in range(self.world) for _, _, ln in self.transfers(kk, q, d))

    def buffer_elems(self, q: int):
        """(A, B): complex elements of the two ping-pong layer buffers of rank q (layers n-1, n-3, .. in A; n-2, n-4, .. in B);
        the replicated layers below k0 are held whole in rank order in the same buffers"""
        n = self.n
        size = {k: (self.storage(k, q)[1] if k >= self.k0 - 1 else self.count[k]) for k in range(0, n)}
        size[self.k0 - 1] = max(size.get(self.k0 - 1, 1), self.count[self.k0 - 1])
        a = max([size[k] for k in range(n - 1, -1, -2)] + [1])
        b = max([size[k] for k in range(n - 2, -1, -2)] + [1])
        return a, b

    def rho_ranges(self, k: int, q: int):
        """[(lo, hi)] per prefix weight w = 0..k: the prefixes of slab w that rank q computes at layer k ((0, 0): none)"""
        rr = [(0, 0)] * (k + 1)
        for w, a, b in self.own[q]:
            if w <= k:
                assert rr[w] == (0, 0)
                rr[w] = (a, b)
        return rr


class SlabChain:
    """One SLOS chain under SlabPlan on this rank (one process per GPU, torch.distributed).  The compute step is injected so
    that the ownership / exchange logic runs on CPU with gloo in the tests:

      full_fn(k, mk, parent_rank_order, out)                       replicated layer k < k0 in FSArray rank order
      to_slab_fn(k, rank_order, out_slab)                          layer k0-1: rank order -> slab-major (whole layer)
      slab_fn(k, mk, parent_slab, rho_ranges, parent_off, child_off, child, probs, psum)
                                                                   this rank's prefixes of layer k, slab-major in and out

    Layers k0-1 .. n-1 live in two COMPACT slab-major ping-pong buffers (SlabPlan.storage: per slab the prefixes this rank owns
    or reads, nothing else -- 14 photons / 28 modes needs 60 GB of the 192 GB layer 13 on the busiest of 8 ranks); the output
    probabilities are stored compactly too, own slabs back to back (`out_slices`).  After layer k every rank sends, per
    consumer, the rows of its slabs that the consumer's children of layer k+1 read through prefix modes: one NCCL
    send / recv group per layer, contiguous slices only.
    """

    def __init__(self, plan: SlabPlan, order, alloc, alloc_real, full_fn, to_slab_fn, slab_fn, rank: int, group=None):
        import torch
        self.plan, self.order, self.rank, self.group = plan, order, rank, group
        L = plan.layout
        n = plan.n
        self.n = n
        self.full_fn, self.to_slab_fn, self.slab_fn = full_fn, to_slab_fn, slab_fn
        ea, eb = plan.buffer_elems(rank)
        self.buf_a = alloc(ea)
        self.buf_b = alloc(eb)
        # compact output: own slabs of layer n back to back; child_off[w] such that index = off + rho * S + t
        self.out_slices = []           # (w, rho_lo, rho_hi, offset, length)
        self.out_off = [0] * (n + 1)
        acc = 0
        for w, a, b in plan.own[rank]:
            ln = (b - a) * L.S[n][w]
            self.out_slices.append((w, a, b, acc, ln))
            self.out_off[w] = acc - a * L.S[n][w]
            acc += ln
        self.probs = alloc_real(max(acc, 1))[:acc]
        self.psum = torch.zeros(1, dtype=torch.float64, device=self.probs.device)
        self._xfer = {}
        for k in range(plan.k0, n):
            for g in range(plan.groups()):
                sends = [(q, lo, ln) for q in range(plan.world) for lo, _, ln in plan.transfers(k, rank, q, g)]
                recvs = [(q, lo, ln) for q in range(plan.world) for _, lo, ln in plan.transfers(k, q, rank, g)]
                self._xfer[(k, g)] = (sends, recvs)
        self.bytes_received = 16 * plan.recv_elems(rank)
        self.bytes_sent = 16 * plan.send_elems(rank)
        self.bytes = 16 * (ea + eb) + 8 * acc

    def _buf(self, k: int):
        return self.buf_a if (self.n - 1 - k) % 2 == 0 else self.buf_b

    def _scratch(self, elems: int):
        if getattr(self, "_scratch_buf", None) is None or self._scratch_buf.numel() < elems:
            self._scratch_buf = self.buf_a.new_empty(elems)
        return self._scratch_buf[:elems]

    def _exchange(self, k: int, g: int, buf):
        import torch
        import torch.distributed as dist
        sends, recvs = self._xfer[(k, g)]
        if not sends and not recvs:
            return []
        flat = torch.view_as_real(buf)
        ops = [dist.P2POp(dist.irecv, flat[lo:lo + ln], q, self.group) for q, lo, ln in recvs]
        ops += [dist.P2POp(dist.isend, flat[lo:lo + ln], q, self.group) for q, lo, ln in sends]
        return dist.batch_isend_irecv(ops)

    def run(self, reduce_sum: bool = True, on_last=None, mark=None):
        """one step: returns (compact probabilities of this rank's slabs, sum(p)).  ``mark(label)`` is called at the phase
        boundaries of the step (replicated layers, each layer's launches, each wait for a halo) -- bench.py records a CUDA
        event there to print where a step spends its time."""
        import torch.distributed as dist
        plan, n, r = self.plan, self.n, self.rank
        mark = mark or (lambda label: None)
        cnt = plan.count
        self.psum.zero_()
        k0 = plan.k0
        # replicated layers 1 .. k0-1 in rank order, ping-pong in the same two buffers
        parent = None
        for k in range(1, k0):
            buf = self._buf(k)
            self.full_fn(k, self.order[k - 1], parent, buf[:cnt[k]])
            parent = buf[:cnt[k]]
        if k0 - 1 >= 1:
            # layer k0-1 to slab-major: permute through the other ping-pong buffer (free until layer k0 is written)
            other = self._buf(k0)
            scratch = other[:cnt[k0 - 1]] if other.numel() >= cnt[k0 - 1] else self._scratch(cnt[k0 - 1])
            self.to_slab_fn(k0 - 1, parent, scratch)
            self._buf(k0 - 1)[:cnt[k0 - 1]].copy_(scratch)
            parent = self._buf(k0 - 1)
        else:
            parent = self._buf(0)
            parent[:1] = 1.0               # the vacuum: one prefix of weight 0, one tail
        groups = []                    # exchange groups that followed layer k-1, in the order the pieces consume them
        mark("replicated")
        for k in range(k0, n + 1):
            poff = plan.offsets(k - 1, r)
            for g, pc in enumerate(plan.piece[r]):
                rr = [(0, 0)] * (k + 1)
                for w, a, b in pc:
                    if w <= k:
                        assert rr[w] == (0, 0)
                        rr[w] = (a, b)
                if g < len(groups) and groups[g] and self._xfer[(k - 1, g)][1]:
                    for w_ in groups[g]:           # the rows this piece reads have arrived (earlier groups were waited before);
                        w_.wait()                  # a group in which this rank only SENDS is waited for at the end of the layer
                    groups[g] = []
                    mark(f"wait{k}.{g}")
                if not any(hi > lo for lo, hi in rr):
                    continue
                if k < n:
                    self.slab_fn(k, self.order[k - 1], parent, rr, poff, plan.offsets(k, r), self._buf(k), None, None)
                else:
                    if on_last is not None:
                        on_last("begin")
                    self.slab_fn(k, self.order[k - 1], parent, rr, poff, self.out_off, None, self.probs, self.psum)
                    if on_last is not None:
                        on_last("end")
            mark(f"layer{k}")
            for grp in groups:             # sends of layer k-1 read the buffer that layer k+1 is about to overwrite
                for w_ in grp:
                    w_.wait()
            if k < n:
                buf = self._buf(k)
                groups = [self._exchange(k, g, buf) for g in range(plan.groups())]
                parent = buf
            else:
                groups = []
        if reduce_sum and plan.world > 1:
            dist.all_reduce(self.psum, op=dist.ReduceOp.SUM, group=self.group)
        return self.probs, self.psum


def engine_slab_chain(engine, U_ref, in_state, group=None, shard_min: int = 1 << 23, pieces: int = 4):
    """Device instantiation of SlabChain: kernels from libfock_b200.so (slos_layer, slos_layer_slab), NCCL send / recv."""
    import torch
    import torch.distributed as dist
    from .engine import prodnfact
    occ = [int(x) for x in in_state]
    m, n = len(occ), sum(occ)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    plan = SlabPlan(m, n, world, shard_min=shard_min, pieces=pieces)
    L = plan.layout
    order = engine.slos_order(occ)
    inf = prodnfact(occ)
    vac = torch.ones(1, dtype=torch.complex128, device=engine.device)
    perm_cache = {}

    def alloc(k):
        return torch.empty(k, dtype=torch.complex128, device=engine.device)

    def alloc_real(k):
        return torch.empty(k, dtype=torch.float64, device=engine.device)

    def full_fn(k, mk, parent, out):
        engine.slos_layer(m, k, U_ref[0], mk, vac if parent is None else parent, child=out)

    def to_slab_fn(k, src, dst):
        if k not in perm_cache:
            perm_cache[k] = torch.from_numpy(L.permutation(k)).to(engine.device)
        dst.index_copy_(0, perm_cache[k], src)

    def slab_fn(k, mk, parent, rr, poff, coff, child, probs, psum):
        engine.slos_layer_slab(m, k, L.p, U_ref[0], mk, parent, rr, poff, coff, child=child, probs=probs, psum=psum, in_prodnfact=inf)

    return SlabChain(plan, order, alloc, alloc_real, full_fn, to_slab_fn, slab_fn, rank, group)
