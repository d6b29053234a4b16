"""FSArray equivalent on top of the C ABI (count / rank / unrank; reference perceval/utils/states.py:255-298)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def count(m: int, n: int) -> int:
    return _lib.count(m, n)


def rank_states(m: int, n: int, states) -> np.ndarray:
    """Ranks (FSArray.find) of a (count, m) uint8 array; states with a photon number != n map to 2**64-1 (npos)."""
    st = np.ascontiguousarray(np.asarray(states, dtype=np.uint8).reshape(-1, m))
    out = np.empty(st.shape[0], dtype=np.uint64)
    _lib.check(_lib.load().fock_rank_host(m, n, st.ctypes.data_as(C.c_void_p), st.shape[0], out.ctypes.data_as(C.c_void_p)),
               "fock_rank_host")
    return out


def unrank(m: int, n: int, ranks) -> np.ndarray:
    rk = np.ascontiguousarray(np.asarray(ranks, dtype=np.uint64).reshape(-1))
    out = np.empty((rk.shape[0], m), dtype=np.uint8)
    _lib.check(_lib.load().fock_unrank_host(m, n, rk.ctypes.data_as(C.c_void_p), rk.shape[0], out.ctypes.data_as(C.c_void_p)),
               "fock_unrank_host")
    return out


def enumerate_states(m: int, n: int) -> np.ndarray:
    """All states of FSArray(m, n) in order as a (count, m) uint8 array."""
    return unrank(m, n, np.arange(count(m, n), dtype=np.uint64))


def iterate_states(m: int, n: int, chunk: int = 1 << 16):
    """Occupation tuples of FSArray(m, n) in order (descending lexicographic)."""
    total = count(m, n)
    for lo in range(0, total, chunk):
        block = unrank(m, n, np.arange(lo, min(lo + chunk, total), dtype=np.uint64))
        for row in block:
            yield tuple(int(x) for x in row)


NPOS = (1 << 64) - 1
