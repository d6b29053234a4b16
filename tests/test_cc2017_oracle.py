"""CPU: the oracle's Clifford & Clifford Algorithm A is pinned by exact enumeration -- the law of its output equals the
SLOS distribution -- and by statistics of its Philox-driven sampler."""
import numpy as np
import pytest


@pytest.mark.parametrize("m,in_state", [(4, (1, 1, 1, 0)), (4, (2, 1, 0, 0)), (5, (1, 0, 1, 0, 1)), (3, (0, 2, 0)), (6, (1, 1, 0, 0, 1, 1))])
def test_exact_pmf_equals_slos(oracle, m, in_state):
    u = oracle.random_unitary(m, seed=5)
    pmf = oracle.cc2017_exact_pmf(u, in_state)
    ref = oracle.slos_probs(u, in_state, scatter=True)
    assert np.abs(pmf - ref).max() < 5e-15
    assert abs(pmf.sum() - 1) < 1e-13


def test_sampler_statistics_and_support(oracle):
    m, in_state, count = 5, (1, 1, 0, 1, 0), 60000
    u = oracle.random_unitary(m, seed=9)
    smp = oracle.cc2017_samples(u, in_state, count, seed=3)
    assert (smp.sum(axis=1) == 3).all()
    ranks = oracle.rank_batch(m, 3, smp)
    freq = np.bincount(ranks.astype(np.int64), minlength=oracle.count(m, 3)) / count
    p = oracle.slos_probs(u, in_state)
    assert 0.5 * np.abs(freq - p).sum() < 0.02
    # stream is keyed by the sample index: splitting a batch does not change it
    a = oracle.cc2017_samples(u, in_state, 100, seed=3, offset=0)
    b = oracle.cc2017_samples(u, in_state, 50, seed=3, offset=50)
    assert (smp[:100] == a).all() and (a[50:] == b).all()


def test_reference_sampling_pins(oracle):
    # tests/backends/test_backends.py:58-67 and tests/components/test_processor.py:130-162 (HOM: no |1,1>)
    smp = oracle.cc2017_samples(oracle.bs_h(), (0, 1), 10000, seed=1)
    c01 = ((smp[:, 0] == 0) & (smp[:, 1] == 1)).sum()
    assert 4750 < c01 < 5250
    smp = oracle.cc2017_samples(oracle.bs_rx(), (1, 1), 500, seed=2)
    assert not ((smp[:, 0] == 1) & (smp[:, 1] == 1)).any()


def test_philox_uniform_range(oracle):
    xs = np.array([oracle.uniform(7, i, d) for i in range(200) for d in range(5)])
    assert xs.min() >= 0 and xs.max() < 1 and abs(xs.mean() - 0.5) < 0.05
    assert oracle.uniform(7, 3, 2) == oracle.uniform(7, 3, 2) != oracle.uniform(8, 3, 2)
