"""Result wire formats: number format against golden vectors made with the reference's own
perceval/utils/format.py (tests/golden/make_format_golden.py), BSSamples / BSCount / BSDistribution text against the
reference algorithms restated literally from perceval/serialization/_state_serialization.py:68-92 and
perceval/utils/conversion.py:52-70."""
import json
import os
from collections import Counter

import numpy as np
import pytest
import torch

from perceval_b200 import wire

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simple_float.json")


def ref_serialize_bssamples(rows):
    # _state_serialization.py:68-78, literally (bss = list of tuples)
    order = [0] * len(rows)
    mapping = {}
    index = 0
    for idx, bs in enumerate(rows):
        if bs not in mapping:
            mapping[bs] = index
            order[idx] = index
            index += 1
        else:
            order[idx] = mapping[bs]
    return ';'.join(["|" + ",".join(map(str, bs)) + ">" for bs in mapping.keys()]) + '/' + ';'.join([str(i) for i in order])


def random_samples(count, m, n, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros((count, m), dtype=np.uint8)
    for i in range(count):
        for j in rng.integers(0, min(m, 4), n):     # few distinct states -> many repeats
            out[i, j] += 1
    return out


def test_format_probability_matches_reference_simple_float():
    vals = json.load(open(GOLDEN))
    assert len(vals) > 400
    for v, s in vals:
        assert wire.format_probability(v) == s, (v, s, wire.format_probability(v))


@pytest.mark.parametrize("count,m,n", [(1, 3, 2), (257, 6, 3), (5000, 12, 4), (64, 400, 5)])
def test_bssamples_text_cpu(count, m, n):
    smp = random_samples(count, m, n, seed=count)
    rows = [tuple(int(x) for x in r) for r in smp]
    text = wire.serialize_bssamples(torch.from_numpy(smp))
    assert text == ref_serialize_bssamples(rows)
    assert (wire.deserialize_bssamples(text) == smp).all()
    states, counts = wire.samples_to_sample_count(torch.from_numpy(smp))
    ref = Counter(rows)                                   # conversion.py:52-60
    assert [tuple(int(x) for x in r) for r in states.numpy()] == list(ref.keys())
    assert counts.tolist() == list(ref.values())
    st2, probs = wire.samples_to_probs(torch.from_numpy(smp))
    assert abs(float(probs.sum()) - 1.0) < 1e-12 and np.allclose(probs.numpy(), np.array(list(ref.values())) / count)
    packed = wire.serialize_samples(torch.from_numpy(smp), compress=True)
    assert packed.startswith(":PCVL:zip:") and wire.decompress(packed) == ":PCVL:BSSamples:" + text


def test_empty_samples_and_count_text():
    e = torch.zeros((0, 4), dtype=torch.uint8)
    assert wire.serialize_bssamples(e) == "/"
    st = torch.tensor([[1, 0, 1], [0, 2, 0]], dtype=torch.uint8)
    assert wire.serialize_count(st, torch.tensor([3, 5]), compress=False) == ":PCVL:BSCount:{|1,0,1>=3;|0,2,0>=5}"
    assert wire.serialize_distribution(st, [0.5, 1 / 3], compress=False) == ":PCVL:BSDistribution:{|1,0,1>=0.5;|0,2,0>=0.333333}"


@pytest.mark.gpu
def test_bssamples_text_from_device_sampler(oracle):
    """The sampler's device tensor -> text, de-duplicated by FSArray rank on the device, equals the reference algorithm
    applied to the same samples on the host; at 20 photons / 400 modes (rank does not fit 64 bits) rows are compared."""
    from perceval_b200.engine import FockEngine
    eng = FockEngine.get(0)
    for m, st, count in [(8, (1, 1, 1, 0, 0, 0, 0, 0), 20000), (400, (1,) * 20 + (0,) * 380, 96)]:
        U = eng.unitary(oracle.random_unitary(m, seed=3))
        smp = eng.cc2017_samples(U, st, count, seed=11)
        rows = [tuple(int(x) for x in r) for r in smp.cpu().numpy()]
        assert wire.serialize_bssamples(smp, eng) == ref_serialize_bssamples(rows)
        states, counts = wire.samples_to_sample_count(smp, eng)
        ref = Counter(rows)
        assert [tuple(int(x) for x in r) for r in states.cpu().numpy()] == list(ref.keys())
        assert counts.cpu().tolist() == list(ref.values())


@pytest.mark.gpu
def test_distribution_text_from_backend(oracle):
    import perceval_b200 as pb
    b = pb.BackendFactory.get_backend("SLOS_B200")
    u = oracle.random_unitary(6, seed=5)
    b.set_circuit(pb.UnitaryCircuit(u))
    b.set_input_state(pb.BasicState([1, 0, 1, 0, 1, 0]))
    text = wire.serialize_backend_distribution(b, compress=False)
    bsd = b.prob_distribution()
    ref = ":PCVL:BSDistribution:{" + ";".join("%s=%s" % (str(k), wire.format_probability(v)) for k, v in bsd.items()) + "}"
    assert text == ref
    assert wire.decompress(wire.serialize_backend_distribution(b)) == ref
