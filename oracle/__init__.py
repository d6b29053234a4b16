"""CPU ORACLE -- test infrastructure, not product code.

ctypes front-end of ``oracle/fock_oracle.c`` (a plain-C restatement of the Fock-amplitude path the reference
delegates to the closed ``exqalibur`` wheel) plus a pure-numpy restatement of the circuit-unitary recipes the
reference's known-answer tests use.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  ``perceval_b200`` never does.

Parity status: SLOS / Naive pinned by the reference's golden vectors (tests/test_oracle_golden.py);
Clifford&Clifford sample sequences "parity unpinned" (closed RNG) -- distribution pinned by exact enumeration.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "fock_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u64, i32, vp, dbl = C.c_uint64, C.c_int, C.c_void_p, C.c_double
        L.orc_binom.restype = u64
        L.orc_binom.argtypes = [i32, i32]
        L.orc_count.restype = u64
        L.orc_count.argtypes = [i32, i32]
        L.orc_rank.restype = u64
        L.orc_rank.argtypes = [i32, i32, vp]
        L.orc_unrank.argtypes = [i32, i32, u64, vp]
        L.orc_rank_batch.argtypes = [i32, i32, vp, u64, vp]
        L.orc_unrank_batch.argtypes = [i32, i32, vp, u64, vp]
        L.orc_enumerate.argtypes = [i32, i32, vp]
        L.orc_prodnfact.restype = dbl
        L.orc_prodnfact.argtypes = [i32, vp]
        L.orc_slos_layer_scatter.argtypes = [i32, i32, vp, i32, vp, vp]
        L.orc_slos_layer_gather.argtypes = [i32, i32, vp, i32, vp, vp, u64, u64]
        L.orc_slos_order.argtypes = [i32, vp, vp]
        L.orc_slos_coefs.restype = i32
        L.orc_slos_coefs.argtypes = [i32, vp, vp, vp, i32]
        L.orc_slos_probs.argtypes = [i32, i32, vp, dbl, vp]
        L.orc_slos_amplitudes.argtypes = [i32, i32, vp, dbl, vp]
        L.orc_glynn_range.argtypes = [i32, vp, u64, u64, vp]
        L.orc_glynn_range_ld.argtypes = [i32, vp, u64, u64, vp]
        L.orc_permanent.argtypes = [i32, vp, vp]
        L.orc_permanent_ryser.argtypes = [i32, vp, vp]
        L.orc_naive_submatrix.argtypes = [i32, i32, vp, vp, vp, vp]
        L.orc_naive_amplitude.argtypes = [i32, vp, vp, vp, vp]
        L.orc_uniform.restype = dbl
        L.orc_uniform.argtypes = [u64, u64, C.c_uint32]
        L.orc_cc2017_sample.argtypes = [i32, i32, vp, vp, u64, u64, vp]
        L.orc_cc2017_samples.argtypes = [i32, i32, vp, vp, u64, u64, u64, vp]
        L.orc_cc2017_exact_pmf.argtypes = [i32, i32, vp, vp, vp]
        L.orc_num_threads.restype = i32
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u(u) -> np.ndarray:
    u = np.ascontiguousarray(np.asarray(u, dtype=np.complex128))
    assert u.ndim == 2 and u.shape[0] == u.shape[1]
    return u


def _s(state) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(list(state), dtype=np.uint8))


# ------------------------------------------------------------------ combinatorics

def count(m: int, n: int) -> int:
    return int(lib().orc_count(m, n))


def rank(state) -> int:
    s = _s(state)
    return int(lib().orc_rank(len(s), int(s.sum()), _p(s)))


def unrank(m: int, n: int, r: int) -> tuple:
    s = np.zeros(m, dtype=np.uint8)
    lib().orc_unrank(m, n, r, _p(s))
    return tuple(int(x) for x in s)


def rank_batch(m: int, n: int, states: np.ndarray) -> np.ndarray:
    states = np.ascontiguousarray(states, dtype=np.uint8).reshape(-1, m)
    out = np.empty(states.shape[0], dtype=np.uint64)
    lib().orc_rank_batch(m, n, _p(states), states.shape[0], _p(out))
    return out


def unrank_batch(m: int, n: int, ranks: np.ndarray) -> np.ndarray:
    ranks = np.ascontiguousarray(ranks, dtype=np.uint64)
    out = np.empty((ranks.shape[0], m), dtype=np.uint8)
    lib().orc_unrank_batch(m, n, _p(ranks), ranks.shape[0], _p(out))
    return out


def enumerate_states(m: int, n: int) -> np.ndarray:
    out = np.empty((count(m, n), m), dtype=np.uint8)
    lib().orc_enumerate(m, n, _p(out))
    return out


def enumerate_states_python(m: int, n: int):
    """Pure-python descending-lex enumeration (independent of the C code; tiny sizes)."""
    if m == 1:
        yield (n,)
        return
    for v in range(n, -1, -1):
        for rest in enumerate_states_python(m - 1, n - v):
            yield (v,) + rest


def prodnfact(state) -> float:
    return float(np.prod([math.factorial(int(x)) for x in state], dtype=np.float64)) if len(state) else 1.0


# ------------------------------------------------------------------ SLOS

def slos_order(state) -> list:
    s = _s(state)
    out = np.zeros(max(int(s.sum()), 1), dtype=np.int32)
    lib().orc_slos_order(len(s), _p(s), _p(out))
    return [int(x) for x in out[: int(s.sum())]]


def slos_layer(m: int, k: int, u, mk: int, parent: np.ndarray, scatter: bool = True) -> np.ndarray:
    u = _u(u)
    parent = np.ascontiguousarray(parent, dtype=np.complex128)
    assert parent.shape[0] == count(m, k - 1)
    child = np.empty(count(m, k), dtype=np.complex128)
    if scatter:
        lib().orc_slos_layer_scatter(m, k, _p(u), mk, _p(parent), _p(child))
    else:
        lib().orc_slos_layer_gather(m, k, _p(u), mk, _p(parent), _p(child), 0, child.shape[0])
    return child


def slos_coefs(u, in_state, scatter: bool = False) -> np.ndarray:
    u = _u(u)
    s = _s(in_state)
    m, n = len(s), int(s.sum())
    assert u.shape[0] == m
    coefs = np.empty(count(m, n), dtype=np.complex128)
    rc = lib().orc_slos_coefs(m, _p(u), _p(s), _p(coefs), 1 if scatter else 0)
    assert rc == 0
    return coefs


def slos_probs(u, in_state, scatter: bool = False) -> np.ndarray:
    s = _s(in_state)
    m, n = len(s), int(s.sum())
    coefs = slos_coefs(u, in_state, scatter)
    probs = np.empty(coefs.shape[0], dtype=np.float64)
    lib().orc_slos_probs(m, n, _p(coefs), prodnfact(s), _p(probs))
    return probs


def slos_amplitudes(u, in_state, scatter: bool = False) -> np.ndarray:
    s = _s(in_state)
    m, n = len(s), int(s.sum())
    coefs = slos_coefs(u, in_state, scatter)
    amps = np.empty(coefs.shape[0], dtype=np.complex128)
    lib().orc_slos_amplitudes(m, n, _p(coefs), prodnfact(s), _p(amps))
    return amps


# ------------------------------------------------------------------ permanents / Naive

def permanent(mat, g0: int | None = None, g1: int | None = None) -> complex:
    mat = _u(mat)
    n = mat.shape[0]
    out = np.zeros(2)
    if g0 is None:
        lib().orc_permanent(n, _p(mat), _p(out))
    else:
        lib().orc_glynn_range(n, _p(mat), g0, g1, _p(out))
    return complex(out[0], out[1])


def permanent_extended(mat, g0: int | None = None, g1: int | None = None) -> complex:
    """The same Glynn walk in x87 extended precision (accuracy arbiter at n >= 28, rounded to double at the end)."""
    mat = _u(mat)
    n = mat.shape[0]
    out = np.zeros(2)
    if g0 is None:
        g0, g1 = 0, 1 << max(n - 1, 0)
    lib().orc_glynn_range_ld(n, _p(mat), g0, g1, _p(out))
    return complex(out[0], out[1])


def permanent_ryser(mat) -> complex:
    mat = _u(mat)
    out = np.zeros(2)
    lib().orc_permanent_ryser(mat.shape[0], _p(mat), _p(out))
    return complex(out[0], out[1])


def naive_submatrix(u, in_state, out_state) -> np.ndarray:
    u = _u(u)
    a, b = _s(in_state), _s(out_state)
    n = int(a.sum())
    mat = np.empty((n, n), dtype=np.complex128)
    lib().orc_naive_submatrix(len(a), n, _p(u), _p(a), _p(b), _p(mat))
    return mat


def naive_amplitude(u, in_state, out_state) -> complex:
    u = _u(u)
    a, b = _s(in_state), _s(out_state)
    out = np.zeros(2)
    lib().orc_naive_amplitude(len(a), _p(u), _p(a), _p(b), _p(out))
    return complex(out[0], out[1])


# ------------------------------------------------------------------ Clifford & Clifford

def uniform(seed: int, idx: int, d: int) -> float:
    return float(lib().orc_uniform(seed, idx, d))


def cc2017_samples(u, in_state, count_: int, seed: int = 0, offset: int = 0) -> np.ndarray:
    u = _u(u)
    s = _s(in_state)
    out = np.empty((count_, len(s)), dtype=np.uint8)
    lib().orc_cc2017_samples(len(s), int(s.sum()), _p(u), _p(s), count_, seed, offset, _p(out))
    return out


def cc2017_exact_pmf(u, in_state) -> np.ndarray:
    u = _u(u)
    s = _s(in_state)
    m, n = len(s), int(s.sum())
    assert n <= 4 and m <= 8
    pmf = np.empty(count(m, n), dtype=np.float64)
    lib().orc_cc2017_exact_pmf(m, n, _p(u), _p(s), _p(pmf))
    return pmf


def num_threads() -> int:
    return int(lib().orc_num_threads())


# ------------------------------------------------------------------ KAT circuit recipes (numpy)
# reference perceval/components/unitary_components.py:142-183 (BS conventions), :262-271 (PS), :475-484 (PERM),
# perceval/components/linear_circuit.py:487-503 (u = cU @ u, later components on the left).

def bs_rx(theta=math.pi / 2):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, 1j * s], [1j * s, c]], dtype=np.complex128)


def bs_h(theta=math.pi / 2):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, s], [s, -c]], dtype=np.complex128)


def bs_ry(theta=math.pi / 2):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def ps(phi):
    return np.array([[np.exp(1j * phi)]], dtype=np.complex128)


def perm(p):
    u = np.zeros((len(p), len(p)), dtype=np.complex128)
    for i, v in enumerate(p):
        u[v, i] = 1
    return u


def r_to_theta(r):
    return 2 * math.acos(math.sqrt(r))


def circuit(m: int, *components) -> np.ndarray:
    """components = (first_mode, unitary) in circuit order"""
    u = np.eye(m, dtype=np.complex128)
    for r0, cu in components:
        big = np.eye(m, dtype=np.complex128)
        k = cu.shape[0]
        big[r0:r0 + k, r0:r0 + k] = cu
        u = big @ u
    return u


def postprocessed_cnot() -> np.ndarray:
    """reference perceval/components/core_catalog/postprocessed_cnot.py:52-57, postprocessed_cz.py:50-58"""
    th13 = r_to_theta(1 / 3)
    cz = [(1, perm([2, 1, 3, 0])),
          (0, bs_h(th13)), (2, bs_h(th13)), (4, bs_h(th13)),
          (1, perm([3, 1, 0, 2]))]
    return circuit(6, (2, bs_h()), *cz, (2, bs_h()))


def random_unitary(m: int, seed: int | None = None) -> np.ndarray:
    """reference perceval/utils/matrix.py:141-173 (randn + 1j randn, QR, q @ diag(sign(real(diag r))))"""
    if seed is not None:
        np.random.seed(seed)
    u = np.random.randn(m, m) + 1j * np.random.randn(m, m)
    q, r = np.linalg.qr(u)
    d = np.sign(np.diagonal(np.real(r)))
    return np.matmul(q, np.diag(d)).astype(np.complex128)
