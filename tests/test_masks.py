"""FSMask semantics: the C ABI's host twin (fock_mask_match_host) and the Python FockMask against the reference's own
known answers (reference tests/utils/test_mask.py:32-45; perceval/backends/_abstract_backends.py:103-146) and against
each other on whole FSArrays; the device kernel (fock_mask_match) against both on a GPU."""
import ctypes as C

import numpy as np
import pytest

import perceval_b200 as pb
from perceval_b200 import _lib, fsarray
from perceval_b200.masks import FockMask


def _host_flags(m, n, mask, states, allow_missing=False):
    L = _lib.load()
    conds = mask.conds_array()
    states = np.ascontiguousarray(states, dtype=np.uint8).reshape(-1, m)
    flags = np.zeros(states.shape[0], dtype=np.uint8)
    rc = L.fock_mask_match_host(m, n, conds.ctypes.data_as(C.c_void_p), conds.shape[0], mask.at_least_bits(), int(allow_missing),
                                states.ctypes.data_as(C.c_void_p), states.shape[0], flags.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return flags.astype(bool)


def test_reference_known_answers():
    # tests/utils/test_mask.py:32-38
    mask = FockMask(6, 4, ["    11"])
    for st, missing, want in [((0, 0, 1, 1, 1, 1), False, True), ((0, 0, 1, 1, 1, 0), False, False), ((0, 0, 1, 1, 1, 0), True, True)]:
        assert mask.match(st, missing) is want
        assert bool(_host_flags(6, 4, mask, [st], missing)[0]) is want
    # tests/utils/test_mask.py:41-45
    mask = FockMask(6, 4, ["   011", "   110"])
    for st, want in [((0, 0, 1, 0, 1, 1), True), ((0, 0, 1, 1, 1, 0), True), ((0, 0, 1, 1, 1, 1), False)]:
        assert mask.match(st) is want
        assert bool(_host_flags(6, 4, mask, [st])[0]) is want


@pytest.mark.parametrize("m,n,masks,at_least", [(6, 2, ["    00"], None), (6, 4, ["*1**0 ", "2    *"], None), (5, 3, ["1 1  "], [0]),
                                                (8, 4, ["  0  : 0"], None), (7, 5, ["*******"], None), (4, 6, ["3*1 "], [0, 2])])
def test_host_twin_matches_python_mask(m, n, masks, at_least):
    mask = FockMask(m, n, masks, at_least)
    states = fsarray.enumerate_states(m, n)
    want = np.array([mask.match(tuple(int(x) for x in s)) for s in states])
    assert (_host_flags(m, n, mask, states) == want).all()
    assert (mask.match_array(states) == want).all()
    want_partial = np.array([mask.match(tuple(int(x) for x in s), True) for s in states])
    assert (_host_flags(m, n, mask, states, True) == want_partial).all()


def test_bad_arguments_are_refused():
    L = _lib.load()
    flags = np.zeros(1, dtype=np.uint8)
    st = np.zeros(4, dtype=np.uint8)
    assert L.fock_mask_match_host(4, 2, None, 1, 0, 0, st.ctypes.data_as(C.c_void_p), 1, flags.ctypes.data_as(C.c_void_p)) != 0
    conds = np.zeros((1, 4), dtype=np.int8)
    assert L.fock_mask_match_host(4, 2, conds.ctypes.data_as(C.c_void_p), 0, 0, 0, st.ctypes.data_as(C.c_void_p), 1,
                                  flags.ctypes.data_as(C.c_void_p)) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,masks,at_least", [(6, 4, ["    11"], None), (12, 6, ["**1*0* ** 2*", "0          1"], None),
                                                (16, 8, ["1              0"], [0]), (10, 5, ["          "], None)])
def test_device_mask_matches_host(m, n, masks, at_least):
    import torch
    from perceval_b200.engine import FockEngine
    eng = FockEngine.get(0)
    mask = FockMask(m, n, masks, at_least)
    states = fsarray.enumerate_states(m, n)
    for missing in (False, True):
        got = eng.mask_flags(m, n, mask, allow_missing=missing).cpu().numpy().astype(bool)
        assert (got == _host_flags(m, n, mask, states, missing)).all()
    N = states.shape[0]
    b, e = N // 3, 2 * N // 3 + 1
    part = eng.mask_flags(m, n, mask, b, e).cpu().numpy().astype(bool)
    assert (part == _host_flags(m, n, mask, states[b:e])).all()
    ranks = eng.mask_ranks(m, n, mask).cpu().numpy()
    assert (ranks == np.nonzero(_host_flags(m, n, mask, states))[0]).all()
    assert torch.cuda.is_available()


@pytest.mark.gpu
def test_slos_backend_masked_outputs_on_device(oracle):
    """SLOS_B200 with a mask: kept states, their order and probabilities equal the unmasked run filtered by the mask
    (the reference's masked distribution is NOT renormalised, docs/source/reference/backends/index.rst:17-18)."""
    from perceval_b200.backends import BackendFactory
    from perceval_b200.circuit import UnitaryCircuit
    from perceval_b200.states import BasicState
    m, st = 8, (1, 1, 0, 1, 0, 1, 0, 0)
    u = oracle.random_unitary(m, seed=3)
    ref = oracle.slos_probs(u, st)
    states = fsarray.enumerate_states(m, 4)
    mask = FockMask(m, 4, ["*0****1*", "2*******"])
    keep = mask.match_array(states)
    b = BackendFactory.get_backend("SLOS_B200")
    b.set_mask(["*0****1*", "2*******"])
    b.set_circuit(UnitaryCircuit(u))
    b.set_input_state(BasicState(list(st)))
    probs = np.array(b.all_prob())
    assert probs.shape[0] == keep.sum()
    assert np.abs(probs - ref[keep]).max() < 1e-12
    it = b._get_iterator(b._input_state)
    assert [tuple(int(x) for x in s) for s in it] == [tuple(int(x) for x in s) for s in states[keep]]
    bsd = b.prob_distribution()
    assert len(bsd) == keep.sum()
    b.clear_mask()
    assert len(b.all_prob()) == states.shape[0]
