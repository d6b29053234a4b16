"""Conformance with a REAL Perceval install: activates only when `import perceval` works (it cannot in
the build image: the exqalibur wheel is not installable offline), then replays the reference's own backend tests
(/root/reference/tests/backends/test_backends.py:39-289) against the registered *_B200 names through Perceval's own
BackendFactory, Circuit components, Processor and Simulator -- the drop-in path of perceval_b200/_compat.py."""
import math

import pytest

pcvl = pytest.importorskip("perceval", reason="Perceval (with exqalibur) is not installed")
torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

STRONG = ["SLOS_B200", "Naive_B200"]


@pytest.fixture(scope="module", autouse=True)
def _register():
    import perceval_b200
    assert perceval_b200.HAVE_PERCEVAL, "perceval imported but perceval_b200 fell back to its mirrors"
    perceval_b200.register()
    from perceval.backends import BACKEND_LIST
    assert all(name in BACKEND_LIST for name in STRONG + ["CliffordClifford2017_B200"])


def check_output_distribution(backend, input_state, expected):
    backend.set_input_state(input_state)
    probs = []
    for output_state, prob in backend.prob_distribution().items():
        want = expected.get(output_state)
        assert pytest.approx(0 if want is None else want) == prob
        probs.append(prob)
    assert pytest.approx(sum(probs)) == 1


def _assert_cnot(backend):
    from perceval.utils import BasicState
    s00, s01 = BasicState([1, 0, 1, 0, 0, 0]), BasicState([1, 0, 0, 1, 0, 0])
    s10, s11 = BasicState([0, 1, 1, 0, 0, 0]), BasicState([0, 1, 0, 1, 0, 0])
    for inp, hit, miss in [(s00, s00, s01), (s01, s01, s00), (s10, s11, s10), (s11, s10, s11)]:
        backend.set_input_state(inp)
        assert pytest.approx(backend.probability(hit)) == 1 / 9
        assert pytest.approx(backend.probability(miss)) == 0


@pytest.mark.parametrize("name", STRONG)
def test_is_a_perceval_backend(name):
    from perceval.backends import AStrongSimulationBackend, BackendFactory
    assert isinstance(BackendFactory.get_backend(name), AStrongSimulationBackend)


@pytest.mark.parametrize("name", STRONG)
def test_wiring_identity_wrong_size(name):
    from perceval.backends import BackendFactory
    from perceval.components import Circuit
    from perceval.utils import BasicState
    b = BackendFactory.get_backend(name)
    b.set_circuit(Circuit(1))
    check_output_distribution(b, BasicState([1]), {BasicState("|1>"): 1})
    b.set_circuit(Circuit(2))
    for s in ([0, 0], [0, 1], [1, 1]):
        check_output_distribution(b, BasicState(s), {BasicState(s): 1})
    with pytest.raises(AssertionError):
        b.set_circuit(Circuit(2))
        b.set_input_state(BasicState([1, 1, 1]))


@pytest.mark.parametrize("name", STRONG)
def test_sym_and_asym_bs(name):
    from perceval.backends import BackendFactory
    from perceval.components import BS
    from perceval.utils import BasicState
    b = BackendFactory.get_backend(name)
    b.set_circuit(BS.H())
    check_output_distribution(b, BasicState("|2,0>"), {BasicState("|2,0>"): 0.25, BasicState("|1,1>"): 0.5, BasicState("|0,2>"): 0.25})
    check_output_distribution(b, BasicState("|1,1>"), {BasicState("|2,0>"): 0.5, BasicState("|0,2>"): 0.5})
    b.set_circuit(BS.H(BS.r_to_theta(1 / 3)))
    check_output_distribution(b, BasicState("|1,0>"), {BasicState("|1,0>"): 1 / 3, BasicState("|0,1>"): 2 / 3})


@pytest.mark.parametrize("name", STRONG)
def test_cnot_and_masks(name):
    from perceval.backends import BackendFactory
    from perceval.components import catalog
    from perceval.utils import BasicState
    cnot = catalog["postprocessed cnot"].build_circuit()
    b = BackendFactory.get_backend(name)
    b.set_circuit(cnot)
    _assert_cnot(b)
    b = BackendFactory.get_backend(name)
    b.set_mask("****00")
    b.set_circuit(cnot)
    logical00 = BasicState([1, 0, 1, 0, 0, 0])
    b.set_input_state(logical00)
    bsd = b.prob_distribution()
    assert len(bsd) == 2 and bsd[logical00] == pytest.approx(1 / 9)
    assert len(b.all_prob()) == len(list(b._get_iterator(logical00)))


def test_slos_refresh_and_evolve():
    from perceval.backends import BackendFactory
    from perceval.components import BS, Circuit
    from perceval.utils import BasicState
    slos = BackendFactory.get_backend("SLOS_B200")
    slos.set_circuit(BS())
    slos.set_input_state(BasicState("|1,1>"))
    slos.set_input_state(BasicState("|8,5>"))
    check_output_distribution(slos, BasicState("|1,1>"), {BasicState("|0,2>"): 0.5, BasicState("|2,0>"): 0.5})
    slos.set_circuit(Circuit(2))
    check_output_distribution(slos, BasicState("|1,1>"), {BasicState("|1,1>"): 1})
    slos.set_circuit(BS.H())
    slos.set_input_state(BasicState([1, 1]))
    sv = slos.evolve()
    assert abs(abs(sv[BasicState([2, 0])]) - math.sqrt(0.5)) < 1e-9


def test_sampler_and_processor_drop_in():
    from perceval.algorithm import Sampler
    from perceval.backends import BackendFactory
    from perceval.components import BS, Processor
    from perceval.utils import BasicState, BSCount
    cliff = BackendFactory.get_backend("CliffordClifford2017_B200")
    cliff.set_circuit(BS.H())
    cliff.set_input_state(BasicState([0, 1]))
    counts = BSCount()
    for s in cliff.samples(10000):
        counts[s] += 1
    assert 4750 < counts[BasicState("|0,1>")] < 5250
    p = Processor("SLOS_B200", BS())
    p.with_input(BasicState([1, 1]))
    probs = Sampler(p).probs()["results"]
    assert probs[BasicState([2, 0])] == pytest.approx(0.5) and probs[BasicState([0, 2])] == pytest.approx(0.5)
