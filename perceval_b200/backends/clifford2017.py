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
    def __init__(self, device=None, seed: int | None = None, prefetch: int = 0):
        """``prefetch``: draw at least this many samples per kernel launch and serve later ``samples()`` calls from the
        device-side pool.  Perceval's sampling loop never asks for more than 1000 samples at once
        (simulators/noisy_sampling_simulator.py:232; SamplesProvider pools <= 2000, :52), far below what fills a B200; since
        sample i of the stream depends on (seed, i) only, the samples handed out are bit-identical with and without a pool."""
        super().__init__()
        self._device = device
        self._engine: FockEngine | None = None
        self._u_dev = None
        self._prefetch = int(prefetch)
        self._pool = None
        self._pool_pos = 0
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
        self._u_dev = self._eng().unitary(self._umat)
        self._pool = None

    def set_input_state(self, input_state):
        super().set_input_state(input_state)
        self._pool = None

    def samples_tensor(self, count: int) -> torch.Tensor:
        """(count, m) uint8 device tensor of occupation numbers: samples [drawn, drawn + count) of the stream."""
        assert self._input_state is not None, "Input state must be set before sampling"
        count = int(count)
        occ = [int(x) for x in self._input_state]
        if self._prefetch <= count:
            if self._pool is not None and self._pool_pos < self._pool.shape[0]:
                head = self._pool[self._pool_pos:self._pool_pos + count]      # pooled samples come first: same stream order
                self._pool_pos += head.shape[0]
                self._drawn += head.shape[0]
                if head.shape[0] == count:
                    return head
                rest = self._eng().cc2017_samples(self._u_dev, occ, count - head.shape[0], self._seed, self._drawn)
                self._drawn += count - head.shape[0]
                return torch.cat([head, rest])
            out = self._eng().cc2017_samples(self._u_dev, occ, count, self._seed, self._drawn)
            self._drawn += count
            return out
        parts, need = [], count
        while need > 0:
            if self._pool is None or self._pool_pos >= self._pool.shape[0]:
                # the pool always starts at the next undrawn index of the stream
                self._pool = self._eng().cc2017_samples(self._u_dev, occ, self._prefetch, self._seed, self._drawn)
                self._pool_pos = 0
            take = self._pool[self._pool_pos:self._pool_pos + need]
            self._pool_pos += take.shape[0]
            self._drawn += take.shape[0]
            need -= take.shape[0]
            parts.append(take)
        return parts[0] if len(parts) == 1 else torch.cat(parts)

    def sample(self):
        return self.samples(1)[0]

    def samples(self, count: int):
        occ = self.samples_tensor(count).cpu().numpy()
        res = BSSamples()
        for row in occ:
            res.append(FockState([int(x) for x in row]))
        return res
