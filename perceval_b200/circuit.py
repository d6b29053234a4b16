"""Smallest circuit object the backends accept: anything with ``m``, ``compute_unitary()`` and
``requires_polarization`` (what ABackend.set_circuit reads, reference perceval/backends/_abstract_backends.py:45-54).
With Perceval installed, pass its own ``Circuit`` / ``Unitary`` objects instead."""
from __future__ import annotations

import numpy as np


class UnitaryCircuit:
    requires_polarization = False

    def __init__(self, u, name: str = "U"):
        """``u``: numpy array (or anything np.asarray accepts) -- or a torch tensor, host (pinned) or device, which the
        backends take as it is (a device tensor makes ``set_circuit`` copy nothing)."""
        if type(u).__module__.startswith("torch"):
            assert u.dim() == 2 and u.shape[0] == u.shape[1], "unitary must be square"
        else:
            u = np.asarray(u, dtype=np.complex128)
            assert u.ndim == 2 and u.shape[0] == u.shape[1], "unitary must be square"
        self._u = u
        self.name = name

    @property
    def m(self) -> int:
        return self._u.shape[0]

    def compute_unitary(self, use_symbolic: bool = False, **kwargs):
        return self._u

    @property
    def U(self):
        return self._u


def random_unitary(m: int, seed: int | None = None) -> np.ndarray:
    """Haar-ish unitary exactly as reference perceval/utils/matrix.py:141-173 (Matrix.random_unitary):
    randn + 1j*randn, QR, q @ diag(sign(real(diag r))); ``seed`` calls np.random.seed first."""
    if seed is not None:
        np.random.seed(seed)
    u = np.random.randn(m, m) + 1j * np.random.randn(m, m)
    q, r = np.linalg.qr(u)
    return np.matmul(q, np.diag(np.sign(np.diagonal(np.real(r))))).astype(np.complex128)
