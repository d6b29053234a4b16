"""Minimal stand-ins for the exqalibur state types the backends exchange with their callers.

When Perceval (and therefore exqalibur) is importable, ``perceval_b200._compat`` hands out the real
``FockState / BSDistribution / StateVector / BSSamples`` instead and nothing in this module is used.  These
classes only implement what the backend contract needs: ``.m``, ``.n``, indexing, iteration,
``prodnfact()``, hashing, ``BSDistribution.add``, ``StateVector +=  state * amplitude``, list-like ``BSSamples``.
"""
from __future__ import annotations

import math
from collections import Counter


class FockState:
    __slots__ = ("_s",)

    def __init__(self, src=()):
        if isinstance(src, FockState):
            self._s = src._s
        elif isinstance(src, str):
            t = src.strip()
            if not (t.startswith("|") and t.endswith(">")):
                raise ValueError(f"cannot parse Fock state {src!r}")
            body = t[1:-1].strip()
            self._s = tuple(int(x) for x in body.split(",")) if body else ()
        elif isinstance(src, int):
            self._s = (0,) * src
        else:
            self._s = tuple(int(x) for x in src)
        if any(x < 0 for x in self._s):
            raise ValueError("negative photon count")

    @property
    def m(self) -> int:
        return len(self._s)

    @property
    def n(self) -> int:
        return sum(self._s)

    def __len__(self):
        return len(self._s)

    def __getitem__(self, i):
        r = self._s[i]
        return FockState(r) if isinstance(i, slice) else r

    def __iter__(self):
        return iter(self._s)

    def __hash__(self):
        return hash(self._s)

    def __eq__(self, other):
        if isinstance(other, FockState):
            return self._s == other._s
        if isinstance(other, (tuple, list)):
            return self._s == tuple(other)
        return NotImplemented

    def __repr__(self):
        return "|" + ",".join(str(x) for x in self._s) + ">"

    __str__ = __repr__

    def prodnfact(self) -> float:
        p = 1
        for x in self._s:
            p *= math.factorial(x)
        return float(p)

    def __mul__(self, other):
        if isinstance(other, FockState):  # tensor product
            return FockState(self._s + other._s)
        if isinstance(other, (int, float, complex)):
            sv = StateVector()
            sv._d[self] = complex(other)
            return sv
        return NotImplemented

    __rmul__ = __mul__

    def threshold_detection(self, nb: int = 1):
        return FockState(min(nb, x) for x in self._s)


BasicState = FockState


class BSDistribution(dict):
    """dict FockState -> probability; ``add`` accumulates and, like the reference distribution type, ignores
    contributions that are not above global_params["min_p"] = 1e-16 (perceval/utils/globals.py:30-34) -- the
    reference's masked-CNOT test (tests/backends/test_backends.py:203-218, ``len(bsd) == 2``) relies on it."""

    MIN_P = 1e-16

    def add(self, state, p: float):
        if p > self.MIN_P:
            self[state] = self.get(state, 0.0) + float(p)

    @property
    def m(self):
        for s in self:
            return s.m
        return 0


class StateVector:
    """Superposition sum_k amp_k |s_k> (un-normalised container; enough for ABackend.evolve)."""

    def __init__(self, src=None):
        self._d: dict[FockState, complex] = {}
        if src is not None:
            self._d[FockState(src)] = 1 + 0j

    def __iadd__(self, other):
        if isinstance(other, FockState):
            other = other * 1
        for s, a in other._d.items():
            v = self._d.get(s, 0j) + a
            self._d[s] = v
        return self

    def __add__(self, other):
        r = StateVector()
        r += self
        r += other
        return r

    def __sub__(self, other):
        return self + other * -1

    def __mul__(self, c):
        r = StateVector()
        for s, a in self._d.items():
            r._d[s] = a * c
        return r

    __rmul__ = __mul__

    def __getitem__(self, s):
        return self._d.get(FockState(s), 0j)

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        return iter(self._d.items())

    def items(self):
        return self._d.items()

    def keys(self):
        return self._d.keys()

    @property
    def m(self):
        for s in self._d:
            return s.m
        return 0

    @property
    def n(self):
        return sorted({s.n for s in self._d})


class BSSamples(list):
    """Chronological list of sampled FockStates."""


class BSCount(Counter):
    pass
