"""Exact parent sets of a contiguous child range of a SLOS layer -- host-side planning for the recompute-window partition.

One SLOS layer is a gather: child s of FSArray(m, k) reads parent s - e_j of FSArray(m, k-1) for every occupied mode j
(reference perceval/backends/_slos.py:91-97, read from the child side).  In FSArray order (descending lexicographic
occupation tuples, reference tests/utils/test_statevector.py:430-438) the map  s -> s - e_j  restricted to the children
with s_j > 0 is an order-preserving bijection onto the whole parent layer.  Hence the parents that the child ranks
[b, e) need through mode j are exactly the parent ranks [C_j(b), C_j(e)), with

    C_j(x) = #{children of rank < x with s_j > 0} = x - #{children of rank < x with s_j = 0},

and the second count has a closed form of the same kind as the rank itself (count the tuples lexicographically greater
than the state of rank x whose j-th entry is 0).  The union over the modes is one or two contiguous ranges in practice;
`parent_segments` merges them (and closes the smallest holes if more than `max_segments` remain).  A rank of a
multi-GPU run keeps only these parents of every layer resident and recomputes them itself -- no exchange step at all
(perceval_b200/dist.py: windowed chain).  Everything here is integer arithmetic on Python ints (ranks reach 3.5e10 at 14
photons / 28 modes).
"""
from __future__ import annotations

from functools import lru_cache
from math import comb as _comb


@lru_cache(maxsize=None)
def comb(a: int, b: int) -> int:
    return _comb(a, b) if a >= 0 and b >= 0 else 0


def count(m: int, n: int) -> int:
    return comb(n + m - 1, n) if m >= 1 else (1 if n == 0 else 0)


def _ways(r: int, modes: int) -> int:
    """states of `modes` modes holding r photons"""
    if modes == 0:
        return 1 if r == 0 else 0
    return comb(r + modes - 1, r)


@lru_cache(maxsize=4096)
def _unrank_cached(m: int, n: int, r: int) -> tuple:
    return tuple(_unrank(m, n, r))


def unrank(m: int, n: int, r: int) -> list:
    return list(_unrank_cached(m, n, r))


def _unrank(m: int, n: int, r: int) -> list:
    s, T = [], n
    for i in range(m - 1):
        q = m - 1 - i
        Tn = T
        while (comb(Tn - 1 + q, q) if Tn > 0 else 0) > r:
            Tn -= 1
        r -= comb(Tn - 1 + q, q) if Tn > 0 else 0
        s.append(T - Tn)
        T = Tn
    s.append(T)
    return s


def rank(m: int, s) -> int:
    r, T = 0, sum(s)
    for i in range(m - 1):
        T -= s[i]
        if T > 0:
            r += comb(T - 1 + (m - 1 - i), m - 1 - i)
    return r


def count_before_with_zero(m: int, k: int, x: int, j: int) -> int:
    """#{states t of FSArray(m, k) with rank < x and t_j == 0}"""
    if x >= count(m, k):
        return _ways(k, m - 1)
    s = _unrank_cached(m, k, x)
    cnt, rem = 0, k
    for i in range(m):
        if i != j and (j > i or s[j] == 0):
            rest = m - i - 1 - (1 if j > i else 0)
            for v in range(s[i] + 1, rem + 1):      # first difference at mode i: t_i = v > s_i, t_j = 0, rest free
                cnt += _ways(rem - v, rest)
        rem -= s[i]
        if i == j and s[j] > 0:
            break                                   # a longer common prefix would need t_j = s_j > 0
    return cnt


def parent_range(m: int, k: int, b: int, e: int, j: int):
    """parent ranks [lo, hi) that the child ranks [b, e) of layer k read through mode j (empty: lo == hi)"""
    return b - count_before_with_zero(m, k, b, j), e - count_before_with_zero(m, k, e, j)


def merge_segments(ranges, max_segments: int = 2):
    """union of half-open ranges as a sorted list; closes the smallest holes until at most max_segments remain"""
    iv = sorted((a, c) for a, c in ranges if c > a)
    out = []
    for a, c in iv:
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], c))
        else:
            out.append((a, c))
    while len(out) > max_segments:
        gaps = [(out[i + 1][0] - out[i][1], i) for i in range(len(out) - 1)]
        _, i = min(gaps)
        out[i:i + 2] = [(out[i][0], out[i + 1][1])]
    return out


def parent_segments(m: int, k: int, child_segments, max_segments: int = 2):
    """parents of layer k-1 needed by the child segments of layer k (k >= 1)"""
    ranges = []
    for b, e in child_segments:
        if e <= b:
            continue
        for j in range(m):
            ranges.append(parent_range(m, k, b, e, j))
    return merge_segments(ranges, max_segments)


def plan_chain(m: int, n: int, b: int, e: int, max_segments: int = 2, full_above: float = 0.8) -> dict:
    """{k: segments of layer k a rank must hold to produce the ranks [b, e) of layer n}, k = n .. 0.
    A layer whose needed part exceeds ``full_above`` of it is taken whole (and with it every layer below): whole layers run
    through the unchecked kernels, which are ~30 % faster per state than the segmented ones, and the deep layers are small."""
    plan = {n: [(b, e)] if e > b else []}
    for k in range(n, 0, -1):
        segs = parent_segments(m, k, plan[k], max_segments)
        if segs and segments_len(segs) >= full_above * count(m, k - 1):
            segs = [(0, count(m, k - 1))]
        plan[k - 1] = segs
    return plan


LAST_LAYER_WEIGHT = 3.0   # measured at 12/24 on 8 ranks: a state of the sharded output layer costs ~37 ps, an inner-layer state ~11 ps


def chain_cost(plan: dict, last_weight: float = LAST_LAYER_WEIGHT) -> float:
    """cost of one plan in inner-layer state updates: every layer it holds, the output range weighted by last_weight"""
    n = max(plan)
    return sum(segments_len(segs) * (last_weight if k == n else 1.0) for k, segs in plan.items() if k >= 1)


@lru_cache(maxsize=64)
def _balanced_boundaries(m: int, n: int, pieces: int, iterations: int, damping: float, last_weight: float) -> tuple:
    """Boundaries x_0 = 0 < ... < x_pieces = N(n) of contiguous output ranges whose chain costs (chain_cost) are about
    equal: equal-count ranges differ by 2.6x in cost at 14 photons / 28 modes because the parents of a range in the
    middle of the layer are spread wider.  Deterministic (every rank computes the same list)."""
    N = count(m, n)
    bounds = [N * i // pieces for i in range(pieces + 1)]
    if pieces <= 1 or N < 4 * pieces:
        return tuple(bounds)
    for _ in range(iterations):
        costs = [max(chain_cost(plan_chain(m, n, bounds[i], bounds[i + 1]), last_weight), 1) for i in range(pieces)]
        target = sum(costs) / pieces
        widths = [(bounds[i + 1] - bounds[i]) * (target / costs[i]) ** damping for i in range(pieces)]
        scale = N / sum(widths)
        acc, new = 0.0, [0]
        for w in widths[:-1]:
            acc += w * scale
            new.append(min(max(int(acc), new[-1] + 1), N - (pieces - len(new))))
        new.append(N)
        bounds = new
    return tuple(bounds)


def balanced_boundaries(m: int, n: int, pieces: int, iterations: int = 8, damping: float = 0.7,
                        last_weight: float = LAST_LAYER_WEIGHT) -> list:
    return list(_balanced_boundaries(m, n, pieces, iterations, damping, float(last_weight)))


def segments_len(segs) -> int:
    return sum(e - b for b, e in segs)


def seg4(segs):
    """the {b0, e0, b1, e1} form of the C ABI (slos_layer_seg)"""
    assert 1 <= len(segs) <= 2
    if len(segs) == 1:
        return [segs[0][0], segs[0][1], segs[0][1], segs[0][1]]
    return [segs[0][0], segs[0][1], segs[1][0], segs[1][1]]
