"""CPU, world_size 2, gloo: the multi-GPU partition / exchange logic of perceval_b200.dist with the compute step
injected from the CPU oracle (the product's device kernels need a GPU; the N>1 plumbing does not)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, exchange, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from perceval_b200 import dist as pdist

    m, in_state = 7, (1, 1, 0, 1, 1, 0, 1)
    n = sum(in_state)
    u = oracle.random_unitary(m, seed=4)
    order = oracle.slos_order(in_state)
    lib = oracle.lib()

    def layer_fn(k, mk, parent, b, e):
        parent_np = np.ones(1, dtype=np.complex128) if parent is None else parent.numpy()
        child = np.empty(e - b, dtype=np.complex128)
        lib.orc_slos_layer_gather(m, k, oracle._p(oracle._u(u)), mk, oracle._p(np.ascontiguousarray(parent_np)), oracle._p(child), b, e)
        return torch.from_numpy(child)

    def last_fn(k, mk, parent, b, e):
        c = layer_fn(k, mk, parent, b, e).numpy()
        states = oracle.unrank_batch(m, k, np.arange(b, e, dtype=np.uint64))
        f = np.array([oracle.prodnfact(s) for s in states])
        p = (np.abs(c) ** 2) * f / oracle.prodnfact(in_state)
        return torch.from_numpy(p), torch.tensor([p.sum()], dtype=torch.float64)

    probs, (b, e), psum, decisions = pdist.slos_probs_sharded(m, in_state, order, oracle.count, layer_fn, last_fn, exchange=exchange)
    ref = oracle.slos_probs(u, in_state)
    assert np.abs(probs.numpy() - ref[b:e]).max() < 1e-14
    assert abs(psum.item() - 1) < 1e-12
    assert all(d == exchange for d in decisions) or exchange == "auto"

    # permanents: batch split (B >= world) and Gray-range split (B < world)
    rng = np.random.default_rng(0)
    mats = torch.from_numpy(rng.standard_normal((5, 6, 6)) + 1j * rng.standard_normal((5, 6, 6)))

    def perm_fn(ms, g0, g1):
        return torch.tensor([oracle.permanent(x.numpy(), g0, g1) if g1 else oracle.permanent(x.numpy()) for x in ms], dtype=torch.complex128)

    want = perm_fn(mats, 0, 0)
    got = pdist.permanents_sharded(perm_fn, mats)
    assert torch.allclose(got, want, rtol=1e-12, atol=0)
    got1 = pdist.permanents_sharded(perm_fn, mats[:1])
    assert torch.allclose(got1, want[:1], rtol=1e-12, atol=0)

    # sampling: disjoint index ranges + gather == the single-process stream
    def sample_fn(cnt, off):
        return torch.from_numpy(oracle.cc2017_samples(u, in_state, cnt, seed=11, offset=off))

    for count in (64, 37):
        allsmp = pdist.samples_sharded(sample_fn, count, offset=5)
        assert (allsmp.numpy() == oracle.cc2017_samples(u, in_state, count, seed=11, offset=5)).all()
    # ragged all-gather helper
    tot = 11
    bb, ee = pdist.shard_range(tot, rank, world)
    full = pdist.all_gather_ragged(torch.arange(bb, ee, dtype=torch.float64), tot)
    assert torch.equal(full, torch.arange(tot, dtype=torch.float64))
    open(os.path.join(outdir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["allgather", "replicate"])
def test_sharded_paths_world2(tmp_path, exchange):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), exchange, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_range_and_exchange_model():
    sys.path.insert(0, ROOT)
    from perceval_b200 import dist as pdist
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [pdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert pdist.choose_exchange(10, 30, 1) == "replicate"
    assert pdist.choose_exchange(286097760, 834451800, 8) in ("replicate", "allgather")


def _windowed_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from perceval_b200 import dist as pdist, partition as P

    m, in_state = 8, (1, 1, 0, 1, 1, 0, 1, 0)
    n = sum(in_state)
    N = P.count(m, n)
    # rank 1 pretends to have very little memory: both ranks must end up with ITS (later) candidate
    free = 10 ** 12 if rank == 0 else 9000   # 0.85 * 9000 = 7650 B: rank 1 only fits from (3, 3.0) on (7344 B)
    sub, weight = pdist.windowed_pick(m, n, rank, world, free, mem_fraction=0.85)
    alone = pdist.windowed_pick(m, n, rank, world, free, mem_fraction=0.85, collective=False)
    picks = [None] * world
    dist.all_gather_object(picks, (sub, weight, alone))
    assert picks[0][:2] == picks[1][:2], picks
    assert picks[0][2] == (1, P.LAST_LAYER_WEIGHT) and picks[1][2] != picks[0][2], picks   # the tight rank forced the choice
    pieces = pdist.windowed_plan(m, n, rank, world, sub, last_weight=weight)
    ranges = [None] * world
    dist.all_gather_object(ranges, (pieces[0][0], pieces[-1][1]))
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == N

    # run the chain of every sub-shard with the oracle's gather layer on packed, segmented buffers
    u = oracle.random_unitary(m, seed=9)
    ref = oracle.slos_probs(u, in_state)
    order = oracle.slos_order(in_state)
    lib = oracle.lib()
    psum = 0.0
    for b, e, plan in pieces:
        layer = {0: np.ones(1, dtype=np.complex128)}
        for k in range(1, n + 1):
            full_parent = np.full(P.count(m, k - 1), np.nan + 0j, dtype=np.complex128)
            off = 0
            for lo, hi in plan[k - 1]:
                full_parent[lo:hi] = layer[k - 1][off:off + hi - lo]
                off += hi - lo
            parts = []
            for lo, hi in plan[k]:
                child = np.empty(hi - lo, dtype=np.complex128)
                lib.orc_slos_layer_gather(m, k, oracle._p(oracle._u(u)), order[k - 1], oracle._p(full_parent), oracle._p(child), lo, hi)
                parts.append(child)
            layer[k] = np.concatenate(parts)
        states = oracle.unrank_batch(m, n, np.arange(b, e, dtype=np.uint64))
        f = np.array([oracle.prodnfact(s) for s in states])
        p = (np.abs(layer[n]) ** 2) * f / oracle.prodnfact(in_state)
        assert np.abs(p - ref[b:e]).max() < 1e-14
        psum += float(p.sum())
    t = torch.tensor([psum], dtype=torch.float64)
    dist.all_reduce(t)
    assert abs(t.item() - 1.0) < 1e-12
    open(os.path.join(outdir, f"wok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world2_windowed_partition_agreement(tmp_path):
    """Recompute-window partition at world size 2 (gloo): the (sub-shards, balance) choice is agreed by all-reduce even when
    the ranks see different free memory, the ranges tile the output layer, every sub-shard chain reproduces its slice and
    the all-reduced sum(p) is 1."""
    world = 2
    port = _free_port()
    mp.spawn(_windowed_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"wok{r}")) for r in range(world))


def _exchange_worker(rank, world, port, outdir, pieces, fractions):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from perceval_b200 import dist as pdist

    m, in_state = 8, (1, 1, 0, 1, 1, 0, 1, 1)
    n = sum(in_state)
    u = oracle.random_unitary(m, seed=10)
    ref = oracle.slos_probs(u, in_state)
    order = oracle.slos_order(in_state)
    lib = oracle.lib()
    plan = pdist.ExchangePlan(m, n, world, pieces=pieces, shard_min=30, fractions=fractions)
    assert 1 <= plan.k0 < n
    # the own ranges tile every sharded layer, and what r receives from q is what q sends to r
    for k in range(plan.k0, n + 1):
        assert plan.own[k][0][0] == 0 and plan.own[k][-1][1] == plan.count[k]
        assert all(a[1] == b[0] for a, b in zip(plan.own[k], plan.own[k][1:]))
    assert sum(plan.recv_elems(r) for r in range(world)) == sum(plan.send_elems(r) for r in range(world))

    def alloc(k):
        return torch.full((k,), float("nan"), dtype=torch.complex128)     # a parent that was never written poisons its children

    def layer_fn(k, mk, parent, out, b, e):
        parent_np = np.ones(1, dtype=np.complex128) if parent is None else parent[:oracle.count(m, k - 1)].numpy()
        child = np.empty(e - b, dtype=np.complex128)
        lib.orc_slos_layer_gather(m, k, oracle._p(oracle._u(u)), mk, oracle._p(np.ascontiguousarray(parent_np)), oracle._p(child), b, e)
        out.copy_(torch.from_numpy(child))

    def last_fn(k, mk, parent, out, psum, b, e):
        c = torch.empty(e - b, dtype=torch.complex128)
        layer_fn(k, mk, parent, c, b, e)
        states = oracle.unrank_batch(m, k, np.arange(b, e, dtype=np.uint64))
        f = np.array([oracle.prodnfact(s) for s in states])
        p = (np.abs(c.numpy()) ** 2) * f / oracle.prodnfact(in_state)
        out.copy_(torch.from_numpy(p))
        psum += float(p.sum())

    chain = pdist.ExchangeChain(m, in_state, order, plan, alloc, layer_fn, last_fn)
    for _ in range(2):                                                       # a second step re-uses buffers and plan
        chain.buf_a.fill_(float("nan"))
        chain.buf_b.fill_(float("nan"))
        probs, (b, e), psum = chain.run()
        assert not torch.isnan(probs).any()
        assert np.abs(probs.numpy() - ref[b:e]).max() < 1e-14
        assert abs(psum.item() - 1.0) < 1e-12
    assert chain.bytes_received == 16 * plan.recv_elems(rank) and chain.bytes_received > 0
    open(os.path.join(outdir, f"xok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,pieces,fractions", [(2, 1, None), (2, 3, None), (3, 4, None), (2, 2, [0.0, 0.3, 1.0])])
def test_exchange_chain_gloo(tmp_path, world, pieces, fractions):
    """Owner-computes + halo exchange at world size 2 / 3 on CPU: every rank computes only its own range of
    every sharded layer, receives exactly the parent slices its children need (NaN-poisoned buffers: a parent that was
    neither owned nor received would surface), and the slices reproduce the oracle's distribution."""
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path), pieces, fractions), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"xok{r}")) for r in range(world))


def _slab_worker(rank, world, port, outdir, m, in_state, shard_min, pieces=4):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from perceval_b200 import slab

    n = sum(in_state)
    u = oracle.random_unitary(m, seed=12)
    ref = oracle.slos_probs(u, in_state)
    order = oracle.slos_order(in_state)
    lib = oracle.lib()
    plan = slab.SlabPlan(m, n, world, shard_min=shard_min, pieces=pieces)
    L = plan.layout
    assert 1 <= plan.k0 <= n
    # the pieces of a rank tile its own prefixes exactly once
    for q in range(world):
        got = sorted(x for pc in plan.piece[q] for x in pc)
        merged = []
        for w, a, b in got:
            if merged and merged[-1][0] == w and merged[-1][2] == a:
                merged[-1] = (w, merged[-1][1], b)
            else:
                merged.append((w, a, b))
        assert merged == sorted(plan.own[q]), (merged, plan.own[q])
    # ownership tiles the prefix space of every slab exactly once
    for w in range(n + 1):
        spans = sorted((a, b) for q in range(world) for ww, a, b in plan.own[q] if ww == w)
        assert spans[0][0] == 0 and spans[-1][1] == L.nprefix[w] and all(x[1] == y[0] for x, y in zip(spans, spans[1:]))
    assert sum(plan.recv_elems(q) for q in range(world)) == sum(plan.send_elems(q) for q in range(world))

    def gather_layer(k, mk, parent_rank_order):
        child = np.empty(oracle.count(m, k), dtype=np.complex128)
        lib.orc_slos_layer_gather(m, k, oracle._p(oracle._u(u)), mk, oracle._p(np.ascontiguousarray(parent_rank_order)), oracle._p(child), 0, child.shape[0])
        return child

    def full_fn(k, mk, parent, out):
        out.copy_(torch.from_numpy(gather_layer(k, mk, np.ones(1, dtype=np.complex128) if parent is None else parent.numpy())))

    def to_slab_fn(k, src, dst):
        dst.index_copy_(0, torch.from_numpy(L.permutation(k)), src)

    def slab_fn(k, mk, parent, rr, poff, coff, child, probs, psum):
        # emulate the slab kernel with the oracle: compact slab-major parent -> rank order (parents this rank does not hold,
        # or never received, are NaN and poison exactly the children that read them), whole layer in rank order, then only
        # this rank's prefixes are written
        perm_p, perm_c = L.permutation(k - 1), L.permutation(k)
        parent_slab = np.full(oracle.count(m, k - 1), np.nan + 0j, dtype=np.complex128)
        store, _size = plan.storage(k - 1, rank)
        for w, (lo, hi, base) in enumerate(store):
            S = L.S[k - 1][w]
            assert poff[w] == base - lo * S
            parent_slab[L.off[k - 1][w] + lo * S: L.off[k - 1][w] + hi * S] = parent.numpy()[base: base + (hi - lo) * S]
        child_ro = gather_layer(k, mk, parent_slab[perm_p])
        child_slab = np.empty_like(child_ro)
        child_slab[perm_c] = child_ro
        states_slab = None
        for w, (lo, hi) in enumerate(rr):
            if hi <= lo:
                continue
            S = L.S[k][w]
            src = child_slab[L.off[k][w] + lo * S: L.off[k][w] + hi * S]
            dst0 = coff[w] + lo * S
            if child is not None:
                child[dst0:dst0 + (hi - lo) * S] = torch.from_numpy(src)
            if probs is not None:
                if states_slab is None:
                    st = oracle.enumerate_states(m, k)
                    states_slab = np.empty_like(st)
                    states_slab[perm_c] = st
                ss = states_slab[L.off[k][w] + lo * S: L.off[k][w] + hi * S]
                f = np.array([oracle.prodnfact(x) for x in ss])
                p = (np.abs(src) ** 2) * f / oracle.prodnfact(in_state)
                probs[dst0:dst0 + (hi - lo) * S] = torch.from_numpy(p)
                psum += float(p.sum())

    chain = slab.SlabChain(plan, order, lambda k: torch.full((k,), float("nan"), dtype=torch.complex128),
                           lambda k: torch.full((k,), float("nan"), dtype=torch.float64), full_fn, to_slab_fn, slab_fn, rank)
    ref_slab = np.empty_like(ref)
    ref_slab[L.permutation(n)] = ref
    for _ in range(2):
        chain.buf_a.fill_(float("nan"))
        chain.buf_b.fill_(float("nan"))
        probs, psum = chain.run()
        assert not torch.isnan(probs).any()
        for w, a, b, off, ln in chain.out_slices:
            S = L.S[n][w]
            want = ref_slab[L.off[n][w] + a * S: L.off[n][w] + b * S]
            assert np.abs(probs.numpy()[off:off + ln] - want).max() < 1e-14
        assert abs(psum.item() - 1.0) < 1e-12
    got = torch.tensor([float(probs.numel())])
    dist.all_reduce(got)
    assert int(got.item()) == ref.shape[0]
    open(os.path.join(outdir, f"sok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m,in_state,shard_min,pieces", [(2, 8, (1, 1, 0, 1, 1, 0, 1, 1), 30, 4), (3, 8, (1, 1, 0, 1, 1, 0, 1, 1), 30, 3),
                                                               (2, 6, (2, 0, 1, 1, 0, 1), 5, 1), (2, 8, (1, 1, 1, 1, 0, 0, 0, 0), 1 << 40, 2),
                                                               (3, 12, (1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0), 20, 5)])
def test_slab_chain_gloo(tmp_path, world, m, in_state, shard_min, pieces):
    """Slab partition at world size 2 / 3 on CPU: fixed prefix ownership, tail parents local, prefix rows
    exchanged as contiguous slab slices; NaN-poisoned buffers prove that every parent a rank reads was owned or received."""
    mp.spawn(_slab_worker, args=(world, _free_port(), str(tmp_path), m, in_state, shard_min, pieces), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"sok{r}")) for r in range(world))
