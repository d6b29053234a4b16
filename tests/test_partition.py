"""Host-side planning of the recompute-window partition (perceval_b200/partition.py): exact parent sets of contiguous
child ranges, against brute-force enumeration, and the chain executed with the CPU oracle as the compute step."""
import random

import numpy as np
import pytest

from perceval_b200 import partition as P


def _brute_parents(m, k, b, e):
    need = set()
    for r in range(b, e):
        s = P.unrank(m, k, r)
        for j in range(m):
            if s[j] > 0:
                p = list(s)
                p[j] -= 1
                need.add(P.rank(m, p))
    return need


@pytest.mark.parametrize("m,k", [(3, 2), (4, 5), (6, 4), (5, 7), (8, 3), (7, 5), (2, 6), (1, 3)])
def test_parent_segments_are_exact(m, k):
    rng = random.Random(m * 100 + k)
    N = P.count(m, k)
    for _ in range(25):
        b = rng.randrange(N)
        e = rng.randrange(b + 1, N + 1)
        need = _brute_parents(m, k, b, e)
        exact = P.parent_segments(m, k, [(b, e)], max_segments=64)
        assert set(x for lo, hi in exact for x in range(lo, hi)) == need
        two = P.parent_segments(m, k, [(b, e)], max_segments=2)
        assert len(two) <= 2 and need <= set(x for lo, hi in two for x in range(lo, hi))
        for j in range(m):   # per mode: one contiguous range (the monotone bijection)
            lo, hi = P.parent_range(m, k, b, e, j)
            got = set()
            for r in range(b, e):
                s = P.unrank(m, k, r)
                if s[j] > 0:
                    p = list(s)
                    p[j] -= 1
                    got.add(P.rank(m, p))
            assert got == set(range(lo, hi))


def test_rank_unrank_agree_with_the_library():
    from perceval_b200 import fsarray
    for m, n in [(4, 5), (6, 4), (12, 6)]:
        states = fsarray.enumerate_states(m, n)
        for r in (0, 1, len(states) // 2, len(states) - 1):
            assert P.unrank(m, n, r) == [int(x) for x in states[r]]
            assert P.rank(m, [int(x) for x in states[r]]) == r
        assert P.count(m, n) == len(states)


def test_plan_memory_model_14_28():
    """The flagship configuration fits: 3 sub-shards per rank keep the windowed chain (both layer buffers + the rank's
    probabilities) under 150 GB per GPU on 8 GPUs."""
    from perceval_b200 import dist as pdist
    m, n = 28, 14
    worst = max(pdist.windowed_peak_bytes(n, pdist.windowed_plan(m, n, r, 8, 3)) for r in range(8))
    assert worst < 150e9
    pieces = pdist.windowed_plan(m, n, 1, 8, 2)
    assert all(len(plan[k]) <= 2 for _, _, plan in pieces for k in plan)


def test_windowed_chain_with_the_oracle(oracle):
    """The chain executed on packed, segmented layer buffers with the oracle's gather layer: every sub-shard equals its
    slice of the reference distribution."""
    m, st = 7, (1, 1, 0, 2, 0, 1, 0)
    n = sum(st)
    u = oracle.random_unitary(m, seed=8)
    ref = oracle.slos_probs(u, st)
    order = oracle.slos_order(st)
    lib = oracle.lib()
    N = P.count(m, n)
    for world in (2, 3):
        for r in range(world):
            b, e = N * r // world, N * (r + 1) // world
            plan = P.plan_chain(m, n, b, e)
            layer = {0: np.ones(1, dtype=np.complex128)}
            for k in range(1, n + 1):
                full_parent = np.full(P.count(m, k - 1), np.nan + 0j, dtype=np.complex128)   # NaN marks "not resident"
                off = 0
                for lo, hi in plan[k - 1]:
                    full_parent[lo:hi] = layer[k - 1][off:off + hi - lo]
                    off += hi - lo
                pieces = []
                for lo, hi in plan[k]:
                    child = np.empty(hi - lo, dtype=np.complex128)
                    lib.orc_slos_layer_gather(m, k, oracle._p(oracle._u(u)), order[k - 1], oracle._p(full_parent), oracle._p(child), lo, hi)
                    pieces.append(child)
                layer[k] = np.concatenate(pieces)
                assert not np.isnan(layer[k]).any()   # a missing parent would have propagated a NaN
            states = oracle.unrank_batch(m, n, np.arange(b, e, dtype=np.uint64))
            f = np.array([oracle.prodnfact(s) for s in states])
            p = (np.abs(layer[n]) ** 2) * f / oracle.prodnfact(st)
            assert np.abs(p - ref[b:e]).max() < 1e-14


@pytest.mark.parametrize("m,n,pieces", [(7, 5, 6), (12, 6, 8), (24, 12, 8)])
def test_balanced_boundaries(m, n, pieces):
    N = P.count(m, n)
    bounds = P.balanced_boundaries(m, n, pieces)
    assert bounds[0] == 0 and bounds[-1] == N and len(bounds) == pieces + 1
    assert all(a < b for a, b in zip(bounds[:-1], bounds[1:]))
    cost = [P.chain_cost(P.plan_chain(m, n, bounds[i], bounds[i + 1])) for i in range(pieces)]
    equal = [N * i // pieces for i in range(pieces + 1)]
    cost_eq = [P.chain_cost(P.plan_chain(m, n, equal[i], equal[i + 1])) for i in range(pieces)]
    assert max(cost) / (sum(cost) / pieces) <= max(cost_eq) / (sum(cost_eq) / pieces) + 1e-9
    assert bounds == P.balanced_boundaries(m, n, pieces)   # deterministic: every rank derives the same list


def test_slab_plan_at_the_headline_sizes():
    """SlabPlan at BASELINE's multi-GPU configs (host arithmetic only): ownership tiles the prefix space of every slab once,
    what one rank receives is what the others send, the exchange groups of a rank carry its whole halo exactly once, the halo
    of 12 photons / 24 modes stays below 1.6 GB per rank and step, and 14 photons / 28 modes fits 8 x 180 GB."""
    from perceval_b200 import slab
    for m, n, world, pieces in [(24, 12, 2, 1), (24, 12, 4, 1), (24, 12, 8, 1), (28, 14, 8, 4)]:
        pl = slab.SlabPlan(m, n, world, pieces=pieces)
        L = pl.layout
        assert L.p == m - 16 and L.D == 16
        for w in range(n + 1):
            spans = sorted((a, b) for q in range(world) for ww, a, b in pl.own[q] if ww == w)
            assert spans[0][0] == 0 and spans[-1][1] == L.nprefix[w] and all(x[1] == y[0] for x, y in zip(spans, spans[1:]))
        recv = [pl.recv_elems(q) for q in range(world)]
        send = [pl.send_elems(q) for q in range(world)]
        assert sum(recv) == sum(send) and recv[0] == 0          # rank 0 holds the lowest prefix weights: every row it reads is its own
        for q in range(world):
            grouped = sum(ln for k in range(pl.k0, n) for g in range(pl.groups()) for s_ in range(world)
                          for _, _, ln in pl.transfers(k, s_, q, g))
            assert grouped == recv[q]
        # every rank's share of the output layer within a factor 2 of the mean (the cost model trades states for transfer time)
        shares = [pl.own_elems(n, q) / P.count(m, n) for q in range(world)]
        assert abs(sum(shares) - 1.0) < 1e-12 and max(shares) < 2.0 / world
        if (m, n) == (24, 12):
            assert max(recv) * 16 < 1.6e9
        else:
            per_rank = [16 * sum(pl.buffer_elems(q)) + 8 * pl.own_elems(n, q) for q in range(world)]
            assert max(per_rank) < 0.7 * 180e9, max(per_rank)
            assert max(recv) * 16 < 60e9
