"""GPU: the reference's own backend tests (tests/backends/test_backends.py, cited per test) replayed against the device
backends through BackendFactory -- same names, same assertions, circuits rebuilt from the oracle's numpy recipes."""
import math
import os

import numpy as np
import pytest

import perceval_b200 as pb
from perceval_b200 import BackendFactory, BasicState, UnitaryCircuit

pytestmark = pytest.mark.gpu

STRONG = ["SLOS_B200", "Naive_B200"]


def check_output_distribution(backend, input_state, expected):
    # tests/backends/test_backends.py:70-81
    backend.set_input_state(input_state)
    prob_list = []
    for output_state, prob in backend.prob_distribution().items():
        prob_expected = expected.get(output_state)
        if prob_expected is None:
            assert pytest.approx(0) == prob, "cannot find: %s (prob=%f)" % (str(output_state), prob)
        else:
            assert pytest.approx(prob_expected) == prob, "incorrect value for %s: %f/%f" % (str(output_state), prob, prob_expected)
        prob_list.append(prob)
    assert pytest.approx(sum(prob_list)) == 1


def _assert_cnot(backend):
    # tests/backends/test_backends.py:39-55
    s00, s01 = BasicState([1, 0, 1, 0, 0, 0]), BasicState([1, 0, 0, 1, 0, 0])
    s10, s11 = BasicState([0, 1, 1, 0, 0, 0]), BasicState([0, 1, 0, 1, 0, 0])
    backend.set_input_state(s00)
    assert pytest.approx(backend.probability(s00)) == 1 / 9
    assert pytest.approx(backend.probability(s01)) == 0
    backend.set_input_state(s01)
    assert pytest.approx(backend.probability(s01)) == 1 / 9
    assert pytest.approx(backend.probability(s00)) == 0
    backend.set_input_state(s10)
    assert pytest.approx(backend.probability(s11)) == 1 / 9
    assert pytest.approx(backend.probability(s10)) == 0
    backend.set_input_state(s11)
    assert pytest.approx(backend.probability(s11)) == 0
    assert pytest.approx(backend.probability(s10)) == 1 / 9


def test_clifford_bs(oracle):
    # :58-67
    b = BackendFactory.get_backend("CliffordClifford2017_B200")
    b.set_circuit(UnitaryCircuit(oracle.bs_h()))
    b.set_input_state(BasicState([0, 1]))
    counts = pb._compat.BSCount()
    n_samples = 10000
    for s in b.samples(n_samples):
        counts[s] += 1
    assert n_samples * 0.475 < counts[BasicState("|0,1>")] < n_samples * 0.525
    assert n_samples * 0.475 < counts[BasicState("|1,0>")] < n_samples * 0.525
    one = b.sample()
    assert one.n == 1 and one.m == 2


def test_backend_factory_default(oracle):
    # :84-88 (the stand-alone factory's default "SLOS" resolves to the device backend)
    b = BackendFactory.get_backend()
    b.set_circuit(UnitaryCircuit(oracle.bs_h()))
    check_output_distribution(b, BasicState([1, 0]), {BasicState("|1,0>"): 0.5, BasicState("|0,1>"): 0.5})


@pytest.mark.parametrize("backend_name", STRONG)
def test_backend_wiring_and_identity(backend_name):
    # :91-103
    b = BackendFactory.get_backend(backend_name)
    b.set_circuit(UnitaryCircuit(np.eye(1)))
    check_output_distribution(b, BasicState([1]), {BasicState("|1>"): 1})
    b.set_circuit(UnitaryCircuit(np.eye(2)))
    check_output_distribution(b, BasicState([0, 0]), {BasicState("|0,0>"): 1})
    check_output_distribution(b, BasicState([0, 1]), {BasicState("|0,1>"): 1})
    check_output_distribution(b, BasicState([1, 1]), {BasicState("|1,1>"): 1})


@pytest.mark.parametrize("backend_name", STRONG + ["CliffordClifford2017_B200"])
def test_backend_wrong_size(backend_name):
    # :106-113
    b = BackendFactory.get_backend(backend_name)
    with pytest.raises(AssertionError):
        b.set_circuit(UnitaryCircuit(np.eye(2)))
        b.set_input_state(BasicState([1, 1, 1]))


@pytest.mark.parametrize("backend_name", STRONG)
def test_backend_sym_and_asym_bs(backend_name, oracle):
    # :116-142
    b = BackendFactory.get_backend(backend_name)
    b.set_circuit(UnitaryCircuit(oracle.bs_h()))
    check_output_distribution(b, BasicState("|2,0>"), {BasicState("|2,0>"): 0.25, BasicState("|1,1>"): 0.5, BasicState("|0,2>"): 0.25})
    check_output_distribution(b, BasicState("|1,0>"), {BasicState("|1,0>"): 0.5, BasicState("|0,1>"): 0.5})
    check_output_distribution(b, BasicState("|1,1>"), {BasicState("|2,0>"): 0.5, BasicState("|0,2>"): 0.5})
    b.set_circuit(UnitaryCircuit(oracle.bs_h(2 * math.pi / 3)))
    check_output_distribution(b, BasicState("|2,0>"), {BasicState("|2,0>"): 0.0625, BasicState("|1,1>"): 0.3750, BasicState("|0,2>"): 0.5625})
    check_output_distribution(b, BasicState("|1,0>"), {BasicState("|1,0>"): 0.25, BasicState("|0,1>"): 0.75})


@pytest.mark.parametrize("backend_name", STRONG)
def test_backend_cnot(backend_name, oracle):
    # :170-185
    b = BackendFactory.get_backend(backend_name)
    b.set_circuit(UnitaryCircuit(oracle.postprocessed_cnot()))
    _assert_cnot(b)
    b.set_input_state(BasicState([1, 0, 1, 0, 0, 0]))
    nps = sum(p for s, p in b.prob_distribution().items() if s[4] or s[5])
    assert pytest.approx(nps) == 7 / 9


def test_cnot_with_mask(oracle):
    # :188-200
    b = BackendFactory.get_backend("SLOS_B200")
    b.set_mask(["    00"])
    b.set_circuit(UnitaryCircuit(oracle.postprocessed_cnot()))
    _assert_cnot(b)
    b.set_input_state(BasicState([0, 1, 0, 1, 0, 0]))
    nps = sum(p for s, p in b.prob_distribution().items() if s[4] or s[5])
    assert pytest.approx(nps) == 0


@pytest.mark.parametrize("backend_name", STRONG)
def test_strong_sim_with_mask(backend_name, oracle):
    # :203-218
    b = BackendFactory.get_backend(backend_name)
    b.set_mask("****00")
    b.set_circuit(UnitaryCircuit(oracle.postprocessed_cnot()))
    logical00 = BasicState([1, 0, 1, 0, 0, 0])
    b.set_input_state(logical00)
    bsd = b.prob_distribution()
    assert len(bsd) == 2
    assert bsd[logical00] == pytest.approx(1 / 9)
    assert bsd[BasicState([1, 1, 0, 0, 0, 0])] == pytest.approx(1 / 9)
    assert len(b.all_prob()) == len(list(b._get_iterator(logical00)))


@pytest.mark.parametrize("backend_name", STRONG)
def test_probampli_backends(backend_name, oracle):
    # :221-249
    b = BackendFactory.get_backend(backend_name)
    u = oracle.circuit(3, (0, oracle.bs_h()), (1, oracle.ps(math.pi / 4)), (1, oracle.bs_h()))
    b.set_circuit(UnitaryCircuit(u))
    check_output_distribution(b, BasicState("|0,1,1>"), {
        BasicState("|0,1,1>"): 0, BasicState("|1,1,0>"): 0.25, BasicState("|1,0,1>"): 0.25, BasicState("|2,0,0>"): 0,
        BasicState("|0,2,0>"): 0.25, BasicState("|0,0,2>"): 0.25})
    b.set_circuit(UnitaryCircuit(oracle.bs_rx()))
    check_output_distribution(b, BasicState("|2,3>"), {
        BasicState("|5,0>"): 0.3125, BasicState("|4,1>"): 0.0625, BasicState("|3,2>"): 0.125, BasicState("|2,3>"): 0.125,
        BasicState("|1,4>"): 0.0625, BasicState("|0,5>"): 0.3125})


def test_slos_refresh_coefs(oracle):
    # :252-277 -- same-size circuit change after several inputs must refresh every cached result
    slos = BackendFactory.get_backend("SLOS_B200")
    slos.set_circuit(UnitaryCircuit(oracle.bs_rx()))
    slos.set_input_state(BasicState("|1,1>"))
    slos.set_input_state(BasicState("|8,5>"))
    check_output_distribution(slos, BasicState("|1,1>"), {BasicState("|0,2>"): 0.5, BasicState("|2,0>"): 0.5})
    slos.set_circuit(UnitaryCircuit(np.eye(2)))
    check_output_distribution(slos, BasicState("|1,1>"), {BasicState("|1,1>"): 1})


@pytest.mark.parametrize("backend_name", STRONG)
def test_evolve_indistinguishable(backend_name, oracle):
    # :280-289
    b = BackendFactory.get_backend(backend_name)
    b.set_circuit(UnitaryCircuit(oracle.bs_h()))
    b.set_input_state(BasicState([1, 0]))
    sv = b.evolve()
    assert sv[BasicState([1, 0])] == pytest.approx(math.sqrt(2) / 2) and sv[BasicState([0, 1])] == pytest.approx(math.sqrt(2) / 2)
    b.set_input_state(BasicState([1, 1]))
    sv = b.evolve()
    assert sv[BasicState([2, 0])] == pytest.approx(math.sqrt(2) / 2)
    assert sv[BasicState([0, 2])] == pytest.approx(-math.sqrt(2) / 2)
    assert abs(sv[BasicState([1, 1])]) < 1e-12


def test_iterator_cache_invalidation(oracle):
    # :313-321
    b = BackendFactory.get_backend("Naive_B200")
    u = np.eye(5, dtype=complex)
    u[:2, :2] = oracle.bs_h()
    b.set_circuit(UnitaryCircuit(u))
    b.set_input_state(BasicState([1, 1, 0, 0, 0]))
    b.evolve()
    assert len(b._cache_iterator) != 0
    b.set_circuit(UnitaryCircuit(np.eye(7)))
    assert len(b._cache_iterator) == 0


def test_naive_doc_values_and_mismatch(oracle):
    # docs/source/reference/backends/naive.rst:16-22 ; _naive.py:46-49 special cases
    b = BackendFactory.get_backend("Naive_B200")
    b.set_circuit(UnitaryCircuit(oracle.bs_rx()))
    b.set_input_state(BasicState([1, 1]))
    assert b.prob_amplitude(BasicState([2, 0])) == pytest.approx(0.7071067811865476j)
    assert b.probability(BasicState([1, 1])) == pytest.approx(0)
    assert b.prob_amplitude(BasicState([1, 0])) == 0          # photon-number mismatch
    b.set_input_state(BasicState([1, 0]))
    assert b.prob_amplitude(BasicState([0, 1])) == pytest.approx(1j * math.sqrt(0.5))   # n == 1 shortcut
    b.set_input_state(BasicState([0, 0]))
    assert b.prob_amplitude(BasicState([0, 0])) == 1
    s = BackendFactory.get_backend("SLOS_B200")
    s.set_circuit(UnitaryCircuit(oracle.bs_rx()))
    s.set_input_state(BasicState([1, 1]))
    assert s.prob_amplitude(BasicState([1, 0])) == 0
    assert s.prob_amplitude(BasicState([2, 0])) == pytest.approx(0.7071067811865476j)
    assert b.permanent(np.array([[1, 2], [3, 4]])) == pytest.approx(10)


def test_sampler_golden_value(oracle):
    # tests/algorithm/test_sampler.py:131-158 : 0.38639895265345636
    u = oracle.circuit(2, (0, oracle.bs_rx()), (0, oracle.ps(0.9)), (0, oracle.bs_rx()))
    for name in STRONG:
        b = BackendFactory.get_backend(name)
        b.set_circuit(UnitaryCircuit(u))
        b.set_input_state(BasicState([1, 1]))
        assert b.probability(BasicState([1, 1])) == pytest.approx(0.38639895265345636)


def test_prob_iterator_threshold_and_reiteration(oracle):
    # _abstract_backends.py:197-201 + simulators/_simulator_utils.py:148-171 (walked twice)
    u = oracle.random_unitary(6, seed=2)
    for name in STRONG:
        b = BackendFactory.get_backend(name)
        b.set_circuit(UnitaryCircuit(u))
        b.set_input_state(BasicState([1, 1, 1, 0, 0, 0]))
        it = b.prob_iterator(0.02)
        first = list(it)
        assert first == list(it) and len(first) > 0
        allp = dict(zip(b._get_iterator(b._input_state), b.all_prob()))
        assert {s for s, p in allp.items() if p > 0.02} == {s for s, _ in first}
        for s, p in first:
            assert p == pytest.approx(allp[s], rel=1e-12)


def test_golden_fixtures_permanent_and_sampler():
    from perceval_b200.engine import FockEngine
    import torch
    eng = FockEngine.get(0)
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for n, seed in [(8, 0), (12, 1), (16, 2), (20, 3)]:
        g = np.load(os.path.join(gdir, f"perm_{n}_seed{seed}.npz"))
        got = complex(eng.permanents(torch.from_numpy(g["mat"][None])).cpu().numpy()[0])
        assert abs(got - complex(g["perm"])) <= 1e-10 * abs(complex(g["perm"]))
    g = np.load(os.path.join(gdir, "cc2017_5_10_seed42.npz"))
    got = eng.cc2017_samples(eng.unitary(g["u"]), tuple(g["in_state"]), 256, seed=42, offset=0).cpu().numpy()
    assert (got != g["samples"]).any(axis=1).mean() <= 0.01


def test_bunching_known_answer(oracle):
    # Boson_Bunching.ipynb cell 13: 0.699 % for n = 7
    from tests.test_oracle_golden import bunching_unitary
    n = 7
    u = bunching_unitary(n)
    b = BackendFactory.get_backend("SLOS_B200")
    b.set_circuit(UnitaryCircuit(u))
    b.set_input_state(BasicState([1] * n))
    bunch = sum(b.probability(BasicState([i, n - i] + [0] * (n - 2))) for i in range(n + 1))
    assert round(bunch * 100, 3) == 0.699
