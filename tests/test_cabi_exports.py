"""CPU-only: the C-ABI library loads and exports every symbol include/fock_b200.h declares; host-side helpers work;
compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import perceval_b200 as pb
from perceval_b200 import _lib, fsarray

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fock_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([a-z][a-z0-9_]*)\s*\(", text)
    return sorted({n for n in names if n.startswith(("fock_", "slos_", "glynn_", "naive_", "cc2017_"))})


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.lib_path()) if os.path.exists(_lib.lib_path()) else _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"libfock_b200.so does not export {s}"
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert set(syms) == bound, f"ctypes table and header disagree: {set(syms) ^ bound}"


def test_host_rank_unrank_against_oracle(oracle):
    for m, n in [(1, 0), (1, 3), (2, 5), (3, 2), (4, 5), (6, 4), (12, 6), (5, 0)]:
        N = fsarray.count(m, n)
        assert N == oracle.count(m, n)
        states = fsarray.unrank(m, n, np.arange(N, dtype=np.uint64))
        assert (states == oracle.enumerate_states(m, n)).all()
        assert (fsarray.rank_states(m, n, states) == np.arange(N, dtype=np.uint64)).all()
    # 64-bit ranks
    m, n = 28, 14
    ranks = np.array([0, 1, 2 ** 32 + 12345, 12033222880, oracle.count(m, n) - 1], dtype=np.uint64)
    st = fsarray.unrank(m, n, ranks)
    assert (st == oracle.unrank_batch(m, n, ranks)).all()
    assert (fsarray.rank_states(m, n, st) == ranks).all()
    # photon-number mismatch -> npos (xq.FSArray.find)
    assert int(fsarray.rank_states(3, 2, [[1, 1, 1]])[0]) == fsarray.NPOS


def test_slos_order_matches_reference_rule():
    import ctypes as C
    L = _lib.load()
    for state, exp in [((1, 1, 1, 0), [0, 1, 2]), ((2, 1, 0), [0, 0, 1]), ((0, 3, 1), [1, 1, 1, 2]), ((1, 2), [1, 0, 1])]:
        s = np.array(state, dtype=np.uint8)
        out = np.zeros(8, dtype=np.int32)
        assert L.slos_order(len(s), s.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        assert list(out[:sum(state)]) == exp


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    b = pb.BackendFactory.get_backend("SLOS_B200")
    with pytest.raises(pb.FockError):
        b.set_circuit(pb.UnitaryCircuit(np.eye(2)))
    out = ctypes.c_void_p()
    assert _lib.load().fock_create(0, ctypes.byref(out)) != 0
    assert b"CUDA" in _lib.load().fock_last_error() or b"device" in _lib.load().fock_last_error()


def test_product_never_imports_oracle():
    import subprocess
    import sys
    code = "import sys, perceval_b200, perceval_b200.engine, perceval_b200.dist; assert not any(k == 'oracle' or k.startswith('oracle.') for k in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "perceval_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_backend_registry_and_factory():
    assert {"SLOS_B200", "Naive_B200", "CliffordClifford2017_B200"} <= set(pb.BackendFactory.list())
    assert isinstance(pb.BackendFactory.get_backend("SLOS_B200"), pb.backends.AStrongSimulationBackend)
    assert isinstance(pb.BackendFactory.get_backend("CliffordClifford2017_B200"), pb.backends.ASamplingBackend)
    with pytest.warns(UserWarning):
        assert pb.BackendFactory.get_backend("nope").name == "SLOS_B200"
    assert pb.BackendFactory.get_backend("Naive_B200").name == "Naive_B200"


def test_wrong_size_is_assertion_error():
    # tests/backends/test_backends.py:106-113 -- raised before any device work
    for name in ["Naive_B200", "CliffordClifford2017_B200"]:
        b = pb.BackendFactory.get_backend(name)
        b._circuit = pb.UnitaryCircuit(np.eye(2))  # set_circuit itself needs the device
        with pytest.raises(AssertionError):
            b.set_input_state(pb.BasicState([1, 1, 1]))
    b = pb.BackendFactory.get_backend("SLOS_B200")
    with pytest.raises(AssertionError):
        b.set_input_state(pb.BasicState([1, 1, 1]))  # circuit not set


def test_mask_semantics():
    # tests/utils/test_mask.py:32-45
    from perceval_b200.masks import FockMask
    mask = FockMask(6, 4, ["    11"])
    assert mask.match((0, 0, 1, 1, 1, 1))
    assert not mask.match((0, 0, 1, 1, 1, 0), False)
    assert mask.match((0, 0, 1, 1, 1, 0), True)
    mask = FockMask(6, 4, ["   011", "   110"])
    assert mask.match((0, 0, 1, 0, 1, 1)) and mask.match((0, 0, 1, 1, 1, 0)) and not mask.match((0, 0, 1, 1, 1, 1))
    arr = np.array([(0, 0, 1, 0, 1, 1), (0, 0, 1, 1, 1, 0), (0, 0, 1, 1, 1, 1)], dtype=np.uint8)
    assert mask.match_array(arr).tolist() == [True, True, False]


def test_local_state_types():
    s = pb.BasicState("|1,0,2>")
    assert s.m == 3 and s.n == 3 and s[2] == 2 and list(s) == [1, 0, 2] and s.prodnfact() == 2.0
    assert pb.BasicState([1, 0, 2]) == s and hash(pb.BasicState([1, 0, 2])) == hash(s) and str(s) == "|1,0,2>"
    sv = pb.StateVector()
    sv += s * 0.5j
    sv += s * 0.5j
    assert sv[s] == 1j
    d = pb.BSDistribution()
    d.add(s, 0.25)
    d.add(s, 0.25)
    assert d[s] == 0.5
