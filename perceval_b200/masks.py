"""Host-side FSMask semantics (reference perceval/backends/_abstract_backends.py:103-146, simulators/simulator.py:650-662,
tests/utils/test_mask.py:32-45).

A mask string has one character per mode: ' ' or '*' accepts anything, any other character c fixes the photon count
to ord(c) - 0x30 (digits for 0..9, then ':' ';' ... up to 32).  A state matches the mask set if it matches any mask.
``at_least_modes`` lists modes whose condition means ">= value" instead of "== value".  ``allow_missing=True`` is the
partial match used on intermediate SLOS layers (a state with fewer photons may still grow into a match).

SLOS_B200 applies masks at the output stage: pruned intermediate states can never grow into a kept output, so the kept
amplitudes are identical to the reference's pruned run.
"""
from __future__ import annotations

import numpy as np


class FockMask:
    def __init__(self, m: int, n: int, masks, at_least_modes=None):
        if isinstance(masks, str):
            masks = [masks]
        self.m, self.n = m, n
        self.at_least = set(at_least_modes or [])
        self.conds = []
        for msk in masks:
            assert len(msk) == m, "mask length must equal the number of modes"
            self.conds.append([None if ch in " *" else ord(ch) - 0x30 for ch in msk])

    def match(self, state, allow_missing: bool = False) -> bool:
        occ = list(state)
        for cond in self.conds:
            ok = True
            for i, c in enumerate(cond):
                if c is None:
                    continue
                v = occ[i]
                if allow_missing:
                    if i not in self.at_least and v > c:
                        ok = False
                        break
                elif (v < c) if i in self.at_least else (v != c):
                    ok = False
                    break
            if ok:
                return True
        return False

    def conds_array(self) -> np.ndarray:
        """(nmask, m) int8 for the C ABI (fock_mask_match): -1 accepts anything, v >= 0 fixes the photon count."""
        return np.ascontiguousarray(np.array([[-1 if c is None else int(c) for c in cond] for cond in self.conds], dtype=np.int8))

    def at_least_bits(self) -> int:
        bits = 0
        for i in self.at_least:
            assert 0 <= i < 64
            bits |= 1 << int(i)
        return bits

    def match_array(self, states: np.ndarray) -> np.ndarray:
        """Vectorised exact match over a (count, m) uint8 array -> bool mask."""
        keep = np.zeros(states.shape[0], dtype=bool)
        for cond in self.conds:
            ok = np.ones(states.shape[0], dtype=bool)
            for i, c in enumerate(cond):
                if c is None:
                    continue
                ok &= (states[:, i] >= c) if i in self.at_least else (states[:, i] == c)
            keep |= ok
        return keep
