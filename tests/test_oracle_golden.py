"""Pins the CPU oracle against every golden vector the reference holds for the Fock-amplitude path.

Each test cites the reference test / doc it replays (paths relative to /root/reference).  Nothing here reads
/root/reference at run time: the unitaries are rebuilt with the recipes in oracle/__init__.py.
"""
import itertools
import math

import numpy as np
import pytest

APPROX = dict(rel=1e-6, abs=1e-12)  # pytest.approx default used by the reference tests


def dist(orc, u, in_state, scatter=True):
    in_state = tuple(in_state)
    m, n = len(in_state), sum(in_state)
    p = orc.slos_probs(u, in_state, scatter=scatter)
    states = [tuple(int(x) for x in s) for s in orc.enumerate_states(m, n)]
    return dict(zip(states, p))


def naive_dist(orc, u, in_state):
    in_state = tuple(in_state)
    m, n = len(in_state), sum(in_state)
    out = {}
    for s in orc.enumerate_states(m, n):
        a = orc.naive_amplitude(u, in_state, s)
        out[tuple(int(x) for x in s)] = abs(a) ** 2
    return out


def check(d, expected):
    # check_output_distribution, tests/backends/test_backends.py:70-81
    for s, p in d.items():
        assert p == pytest.approx(expected.get(s, 0), **APPROX), s
    assert sum(d.values()) == pytest.approx(1)


@pytest.fixture(params=["slos_scatter", "slos_gather", "naive"])
def engine(request, oracle):
    if request.param == "slos_scatter":
        return lambda u, s: dist(oracle, u, s, True)
    if request.param == "slos_gather":
        return lambda u, s: dist(oracle, u, s, False)
    return lambda u, s: naive_dist(oracle, u, s)


# ---------------------------------------------------------------- enumeration order pins

def test_order_pins(oracle):
    # tests/utils/test_statevector.py:430-438  max_photon_state_iterator(3, 2)
    exp = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]
    got = []
    for n in range(3):
        got += [tuple(int(x) for x in s) for s in oracle.enumerate_states(3, n)]
    assert got == exp
    # tests/utils/test_density_matrix.py:54-59: |1,1> has index 4 in FockBasis(2,2) = n=0,1,2 concatenated
    basis = []
    for n in range(3):
        basis += [tuple(int(x) for x in s) for s in oracle.enumerate_states(2, n)]
    assert basis.index((1, 1)) == 4
    # tests/utils/test_density_matrix.py:42-52 sizes
    assert oracle.count(3, 12) + sum(oracle.count(3, k) for k in range(12)) == 455
    assert oracle.count(12, 6) == 12376


@pytest.mark.parametrize("m,n", [(1, 0), (1, 3), (2, 5), (3, 2), (4, 5), (6, 4), (12, 6), (5, 0)])
def test_rank_unrank_vs_python_enumeration(oracle, m, n):
    ref = list(oracle.enumerate_states_python(m, n))
    assert len(ref) == oracle.count(m, n) == math.comb(n + m - 1, n)
    got = oracle.enumerate_states(m, n)
    assert [tuple(int(x) for x in s) for s in got] == ref
    ranks = oracle.rank_batch(m, n, got)
    assert (ranks == np.arange(len(ref), dtype=np.uint64)).all()
    assert (oracle.unrank_batch(m, n, ranks) == got).all()


def test_rank_large(oracle):
    # 14 photons / 28 modes needs 64-bit ranks
    m, n = 28, 14
    N = oracle.count(m, n)
    assert N == 35240152720
    for r in [0, 1, N - 1, N // 2, 2 ** 32 + 12345, 12033222880]:
        s = oracle.unrank(m, n, r)
        assert sum(s) == n and oracle.rank(s) == r
    assert oracle.unrank(m, n, 0) == (14,) + (0,) * 27
    assert oracle.unrank(m, n, N - 1) == (0,) * 27 + (14,)


# ---------------------------------------------------------------- known answers, strong simulation

def test_identity(engine, oracle):
    # tests/backends/test_backends.py:97-103 and :91-95
    u = np.eye(2)
    check(engine(u, (0, 0)), {(0, 0): 1})
    check(engine(u, (0, 1)), {(0, 1): 1})
    check(engine(u, (1, 1)), {(1, 1): 1})
    check(engine(np.eye(1), (1,)), {(1,): 1})


def test_sym_bs(engine, oracle):
    # tests/backends/test_backends.py:116-129
    u = oracle.bs_h()
    check(engine(u, (2, 0)), {(2, 0): 0.25, (1, 1): 0.5, (0, 2): 0.25})
    check(engine(u, (1, 0)), {(1, 0): 0.5, (0, 1): 0.5})
    check(engine(u, (1, 1)), {(2, 0): 0.5, (0, 2): 0.5})


def test_asym_bs(engine, oracle):
    # tests/backends/test_backends.py:132-142
    u = oracle.bs_h(2 * math.pi / 3)
    check(engine(u, (2, 0)), {(2, 0): 0.0625, (1, 1): 0.3750, (0, 2): 0.5625})
    check(engine(u, (1, 0)), {(1, 0): 0.25, (0, 1): 0.75})


def test_cnot(engine, oracle):
    # tests/backends/test_backends.py:39-55 and :170-185
    u = oracle.postprocessed_cnot()
    s00, s01, s10, s11 = (1, 0, 1, 0, 0, 0), (1, 0, 0, 1, 0, 0), (0, 1, 1, 0, 0, 0), (0, 1, 0, 1, 0, 0)
    d = engine(u, s00)
    assert d[s00] == pytest.approx(1 / 9) and d[s01] == pytest.approx(0)
    assert sum(p for s, p in d.items() if s[4] or s[5]) == pytest.approx(7 / 9)
    d = engine(u, s01)
    assert d[s01] == pytest.approx(1 / 9) and d[s00] == pytest.approx(0)
    d = engine(u, s10)
    assert d[s11] == pytest.approx(1 / 9) and d[s10] == pytest.approx(0)
    d = engine(u, s11)
    assert d[s11] == pytest.approx(0) and d[s10] == pytest.approx(1 / 9)
    # masked variant :203-218: exactly two outputs of the "****00" sub-space are non-zero for logical 00
    d = engine(u, s00)
    kept = {s: p for s, p in d.items() if s[4] == 0 and s[5] == 0 and p > 1e-12}
    assert len(kept) == 2 and kept[(1, 1, 0, 0, 0, 0)] == pytest.approx(1 / 9)


def test_probampli(engine, oracle):
    # tests/backends/test_backends.py:221-249
    u = oracle.circuit(3, (0, oracle.bs_h()), (1, oracle.ps(math.pi / 4)), (1, oracle.bs_h()))
    check(engine(u, (0, 1, 1)), {(0, 1, 1): 0, (1, 1, 0): 0.25, (1, 0, 1): 0.25, (2, 0, 0): 0, (0, 2, 0): 0.25,
                                 (0, 0, 2): 0.25})
    check(engine(oracle.bs_rx(), (2, 3)), {(5, 0): 0.3125, (4, 1): 0.0625, (3, 2): 0.125, (2, 3): 0.125,
                                           (1, 4): 0.0625, (0, 5): 0.3125})


def test_refresh_case(engine, oracle):
    # tests/backends/test_backends.py:252-277 (values only; the lifecycle is tested on the backend class)
    check(engine(oracle.bs_rx(), (1, 1)), {(0, 2): 0.5, (2, 0): 0.5})
    check(engine(np.eye(2), (1, 1)), {(1, 1): 1})
    d = engine(oracle.bs_rx(), (8, 5))
    assert sum(d.values()) == pytest.approx(1)


def test_evolve_signs(oracle):
    # tests/backends/test_backends.py:280-289 ; Computation_Tutorial.ipynb cell 6 (+-0.7071067811865477)
    u = oracle.bs_h()
    a = oracle.slos_amplitudes(u, (1, 0), scatter=True)
    assert a == pytest.approx([math.sqrt(2) / 2, math.sqrt(2) / 2])
    a = oracle.slos_amplitudes(u, (1, 1), scatter=True)
    assert a == pytest.approx([math.sqrt(2) / 2, 0, -math.sqrt(2) / 2])
    for out, e in zip([(2, 0), (1, 1), (0, 2)], a):
        assert oracle.naive_amplitude(u, (1, 1), out) == pytest.approx(e)


def test_naive_doc(oracle):
    # docs/source/reference/backends/naive.rst:16-22: BS() , |1,0> -> |0,1> amplitude 0.5j*sqrt2... printed values
    u = oracle.bs_rx()
    assert oracle.naive_amplitude(u, (1, 1), (2, 0)) == pytest.approx(0.7071067811865476j)
    # probability 0.5 and the n==1 shortcut of _naive.py:49
    assert oracle.naive_amplitude(u, (1, 0), (0, 1)) == pytest.approx(1j * math.sqrt(0.5))
    assert oracle.naive_amplitude(u, (1, 0), (1, 1)) == 0  # photon-number mismatch -> 0
    assert oracle.naive_amplitude(u, (0, 0), (0, 0)) == 1  # n = 0 -> 1


def test_sampler_golden_value(engine, oracle):
    # tests/algorithm/test_sampler.py:131-158: BS() // PS(phi=0.9) // BS(), |1,1>, two-photon outcomes ->
    # results[|1,1>] == 0.38639895265345636
    u = oracle.circuit(2, (0, oracle.bs_rx()), (0, oracle.ps(0.9)), (0, oracle.bs_rx()))
    assert engine(u, (1, 1))[(1, 1)] == pytest.approx(0.38639895265345636)


def bunching_unitary(n):
    # docs/source/notebooks/Boson_Bunching.ipynb cell 6: rows 0,1 fixed, the rest any orthonormal completion
    w = np.exp(2j * math.pi / (n - 2))
    v1 = np.array([1, 0] + [1 / math.sqrt(2)] * (n - 2), dtype=complex)
    v2 = np.array([0, 1] + [np.conj(w ** (i - 2)) / math.sqrt(2) for i in range(2, n)], dtype=complex)
    basis = [v1 / np.linalg.norm(v1), v2 / np.linalg.norm(v2)]
    rng = np.random.default_rng(1)
    for _ in range(n - 2):
        r = rng.random(n) + 1j * rng.random(n)
        for b in basis:
            r = r - np.vdot(b, r) * b
        basis.append(r / np.linalg.norm(r))
    return np.array(basis)


@pytest.mark.parametrize("n,expected_percent", [(7, 0.699), (8, 0.240)])
def test_boson_bunching_known_answers(oracle, n, expected_percent):
    # Boson_Bunching.ipynb cell 13 output (n=7: 0.699 %) and cell 18 (n=8: 0.240 %): probability that all n
    # indistinguishable photons leave in the first two modes -- the largest known answers in the reference tree.
    u = bunching_unitary(n)
    assert np.abs(u @ u.conj().T - np.eye(n)).max() < 1e-12
    p = oracle.slos_probs(u, (1,) * n)
    assert p.sum() == pytest.approx(1, abs=1e-12)
    bunch = sum(p[oracle.rank((i, n - i) + (0,) * (n - 2))] for i in range(n + 1))
    assert round(bunch * 100, 3) == expected_percent
    bunch_naive = sum(abs(oracle.naive_amplitude(u, (1,) * n, (i, n - i) + (0,) * (n - 2))) ** 2 for i in range(n + 1))
    assert bunch_naive == pytest.approx(bunch, rel=1e-10)


# ---------------------------------------------------------------- cross-engine self-consistency (oracle internals)

@pytest.mark.parametrize("m,in_state", [(5, (1, 1, 1, 1, 0)), (4, (2, 1, 0, 1)), (6, (0, 3, 0, 1, 1, 0)), (3, (0, 0, 4))])
def test_slos_equals_naive_random(oracle, m, in_state):
    u = oracle.random_unitary(m, seed=7)
    a_scatter = oracle.slos_amplitudes(u, in_state, scatter=True)
    a_gather = oracle.slos_amplitudes(u, in_state, scatter=False)
    assert np.abs(a_scatter - a_gather).max() < 1e-14
    n = sum(in_state)
    for r, s in enumerate(oracle.enumerate_states(m, n)):
        assert abs(oracle.naive_amplitude(u, in_state, s) - a_scatter[r]) < 1e-13
    assert abs((np.abs(a_scatter) ** 2).sum() - 1) < 1e-12


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11])
def test_glynn_vs_ryser(oracle, n):
    rng = np.random.default_rng(n)
    mat = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    g, r = oracle.permanent(mat), oracle.permanent_ryser(mat)
    assert abs(g - r) <= 1e-12 * max(1, abs(r))
    # brute force for tiny n
    if n <= 5:
        bf = sum(np.prod([mat[i, p[i]] for i in range(n)]) for p in itertools.permutations(range(n)))
        assert abs(g - bf) <= 1e-12 * max(1, abs(bf))
    # Gray-range split sums to the whole
    if n >= 3:
        tot = 1 << (n - 1)
        parts = oracle.permanent(mat, 0, tot // 3) + oracle.permanent(mat, tot // 3, tot)
        assert abs(parts - g) <= 1e-12 * max(1, abs(g))


def test_slos_order(oracle):
    # _slos.py:61-86: greedy on the mode with most remaining photons, first index on ties
    assert oracle.slos_order((1, 1, 1, 0)) == [0, 1, 2]
    assert oracle.slos_order((2, 1, 0)) == [0, 0, 1]
    assert oracle.slos_order((0, 3, 1)) == [1, 1, 1, 2]
    assert oracle.slos_order((1, 2)) == [1, 0, 1]


def test_glynn_double_vs_extended_precision(oracle):
    """The oracle's double-precision Glynn walk (Kahan sum, column sums re-seeded every 4096 codes) against the same walk in
    x87 extended precision, on Haar sub-matrices whose permanent is orders of magnitude smaller than the terms of the sum."""
    import numpy as np
    for n in (8, 16, 22, 24):
        u = oracle.random_unitary(2 * n, seed=n)
        mat = np.ascontiguousarray(u[:n, :n])
        a, b = oracle.permanent(mat), oracle.permanent_extended(mat)
        assert abs(a - b) <= 2e-11 * abs(b), (n, a, b)
        G = 1 << (n - 1)
        parts = sum(oracle.permanent_extended(mat, q * (G // 4), (q + 1) * (G // 4)) for q in range(4))
        assert abs(parts - b) <= 1e-12 * abs(b)
