"""Naive_B200 -- permanent-based strong simulation on the device.

Mirrors reference perceval/backends/_naive.py:38-71 (NaiveBackend): amplitude = Perm(M) / sqrt(prod(in!) prod(out!)) with
M[r, c] = U[out_mode(r), in_mode(c)]; photon-number mismatch -> 0; n = 0 -> 1; n = 1 -> M[0,0].  The Python triple
loop + xq.permanent_cx of the reference are replaced by the device sub-matrix gather + batched Glynn kernel
(C ABI naive_amplitudes / glynn_permanent_batch); the generic per-output loops of the base class
(_abstract_backends.py:173-208) are replaced by one batched launch over all output states.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import fsarray
from .._compat import MIN_P, AStrongSimulationBackend, BSDistribution, StateVector
from ..engine import FockEngine
from .slos import _StateProbView


class NaiveB200Backend(AStrongSimulationBackend):
    def __init__(self, mask=None, device=None, batch: int = 1 << 16):
        super().__init__()
        self._device = device
        self._engine: FockEngine | None = None
        self._u_dev = None
        self._batch = batch
        if mask is not None:
            self.set_mask(mask)

    @property
    def name(self) -> str:
        return "Naive_B200"

    def _eng(self) -> FockEngine:
        if self._engine is None:
            self._engine = FockEngine.get(self._device)
        return self._engine

    def set_circuit(self, circuit):
        super().set_circuit(circuit)
        self._u_dev = self._eng().unitary(self._umat)

    def prob_amplitude(self, output_state) -> complex:
        istate = self._input_state
        if istate.n != output_state.n:
            return complex(0)
        if istate.n == 0:
            return complex(1)
        st = torch.tensor([[int(x) for x in output_state]], dtype=torch.uint8)
        a = self._eng().naive_amplitudes(self._u_dev, [int(x) for x in istate], out_states=st)
        return complex(a[0].item())

    def permanent(self, matrix) -> complex:
        """xq.permanent_cx equivalent (reference _naive.py:70-71)."""
        m = torch.as_tensor(np.asarray(matrix, dtype=np.complex128))
        return complex(self._eng().permanents(m)[0].item())

    def all_amplitudes_tensor(self, input_state=None) -> torch.Tensor:
        """Amplitudes of every state the iterator yields (mask respected), one batched launch per chunk."""
        if input_state is not None:
            self.set_input_state(input_state)
        ist = self._input_state
        eng = self._eng()
        occ_in = [int(x) for x in ist]
        if self._mask is None and ist.m <= 64:
            total = fsarray.count(ist.m, ist.n)
            outs = []
            for lo in range(0, total, self._batch):
                ranks = torch.arange(lo, min(lo + self._batch, total), dtype=torch.int64, device=eng.device)
                outs.append(eng.naive_amplitudes(self._u_dev, occ_in, out_ranks=ranks))
            return torch.cat(outs) if outs else torch.empty(0, dtype=torch.complex128, device=eng.device)
        states = self._get_iterator(ist)
        arr = torch.tensor([[int(x) for x in s] for s in states], dtype=torch.uint8).view(-1, ist.m)
        outs = [eng.naive_amplitudes(self._u_dev, occ_in, out_states=arr[lo:lo + self._batch])
                for lo in range(0, arr.shape[0], self._batch)]
        return torch.cat(outs) if outs else torch.empty(0, dtype=torch.complex128, device=eng.device)

    def all_prob_tensor(self, input_state=None) -> torch.Tensor:
        a = self.all_amplitudes_tensor(input_state)
        return a.real ** 2 + a.imag ** 2

    def all_prob(self, input_state=None):
        return self.all_prob_tensor(input_state).cpu().tolist()

    def prob_distribution(self):
        probs = self.all_prob_tensor().cpu().tolist()
        bsd = BSDistribution()
        for s, p in zip(self._get_iterator(self._input_state), probs):
            bsd.add(s, p)
        return bsd

    def prob_iterator(self, min_p: float = MIN_P):
        probs = self.all_prob_tensor().cpu().tolist()
        states = self._get_iterator(self._input_state)
        keep = [i for i, p in enumerate(probs) if p > min_p]
        return _StateProbView([states[i] for i in keep], [probs[i] for i in keep])

    def evolve(self):
        amps = self.all_amplitudes_tensor().cpu().numpy()
        res = StateVector()
        for s, a in zip(self._get_iterator(self._input_state), amps):
            res += s * complex(a)
        return res
