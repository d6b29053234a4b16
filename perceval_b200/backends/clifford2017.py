"""CliffordClifford2017_B200 -- exact boson sampling (Clifford & Clifford Algorithm A) on the device.

Mirrors reference perceval/backends/_clifford2017.py:36-61: set_circuit / set_input_state / sample / samples / name.
The sample stream is keyed by (seed, running sample index) with Philox4x32-10, so ``samples(a)`` followed by
``samples(b)`` equals ``samples(a + b)``, and a batch can be sharded over GPUs by index range without changing it.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from .._compat import ASamplingBackend, BSSamples, FockState
from ..engine import FockEngine

_global_seed = None


def set_seed(seed: int | None):
    """Seed used by sampler backends created afterwards (the counterpart of xq.set_seed, perceval/utils/_random.py:55-58)."""
    global _global_seed
    _global_seed = seed


class Clifford2017B200Backend(ASamplingBackend):
    def __init__(self, device=None, seed: int | None = None):
        super().__init__()
        self._device = device
        self._engine: FockEngine | None = None
        self._u_dev = None
        if seed is None:
            seed = _global_seed
        # python's RNG is what pcvl.random_seed() seeds first (perceval/utils/_random.py:36-58)
        self._seed = random.getrandbits(63) if seed is None else int(seed)
        self._drawn = 0

    @property
    def name(self) -> str:
        return "CliffordClifford2017_B200"

    def _eng(self) -> FockEngine:
        if self._engine is None:
            self._engine = FockEngine.get(self._device)
        return self._engine

    def set_circuit(self, circuit):
        super().set_circuit(circuit)
        self._u_dev = self._eng().unitary(np.asarray(self._umat, dtype=np.complex128))

    def set_input_state(self, input_state):
        super().set_input_state(input_state)

    def samples_tensor(self, count: int) -> torch.Tensor:
        """(count, m) uint8 device tensor of occupation numbers."""
        assert self._input_state is not None, "Input state must be set before sampling"
        out = self._eng().cc2017_samples(self._u_dev, [int(x) for x in self._input_state], int(count), self._seed, self._drawn)
        self._drawn += int(count)
        return out

    def sample(self):
        return self.samples(1)[0]

    def samples(self, count: int):
        occ = self.samples_tensor(count).cpu().numpy()
        res = BSSamples()
        for row in occ:
            res.append(FockState([int(x) for x in row]))
        return res
