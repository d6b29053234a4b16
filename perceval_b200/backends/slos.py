"""SLOS_B200 -- device SLOS backend behind Perceval's AStrongSimulationBackend surface.

Mirrors reference perceval/backends/_slos.py:105-223 (SLOSBackend): same constructor keyword (``mask``), same method
names, argument meaning and error behaviour; every number comes from libfock_b200.so (slos_layer / slos_layer_probs /
epilogues).  Differences that are invisible through the ABC contract:
  * only two layers are ever live on the device (the reference keeps all n, _slos.py:44);
  * coefficient vectors are cached per input state on the device (the reference's ``_state_mapping``) and are
    recomputed after a same-size circuit change (_slos.py:136-139, pinned by tests/backends/test_backends.py:252-277);
  * masks are applied at the output stage -- kept amplitudes are identical;
  * tensor-returning variants (``all_prob_tensor``, ``all_amplitudes_tensor``, ``prob_iterator_tensors``) avoid the
    one-Python-object-per-state result types that cannot carry 8e8 states.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import fsarray
from .._compat import MIN_P, AStrongSimulationBackend, BSDistribution, FockState, StateVector
from ..engine import FockEngine, prodnfact


class _StateProbView:
    """Re-iterable (state, probability) pairs (reference _abstract_backends.py:83-90)."""

    def __init__(self, states, probs):
        self.states = states
        self.probs = probs

    def __iter__(self):
        return zip(self.states, self.probs)

    def __len__(self):
        return len(self.probs)


class SLOSB200Backend(AStrongSimulationBackend):
    def __init__(self, mask=None, device=None, max_cached_bytes: int = 16 << 30, use_symbolic: bool = False):
        super().__init__()
        if use_symbolic:
            raise NotImplementedError("SLOS_B200 is numeric (complex128) only; use the reference SLOS backend for sympy")
        self._device = device
        self._engine: FockEngine | None = None
        self._u_dev = None
        self._max_cached_bytes = max_cached_bytes
        self._reset()
        if mask is not None:
            self.set_mask(mask)

    @property
    def name(self) -> str:
        return "SLOS_B200"

    # ------------------------------------------------------------------ lifecycle (_slos.py:117-149)
    def _eng(self) -> FockEngine:
        if self._engine is None:
            self._engine = FockEngine.get(self._device)
        return self._engine

    def _reset(self):
        self._coefs: dict = {}   # input state -> device complex128 coefficients (reference: _state_mapping)
        self._probs: dict = {}   # input state -> device float64 probabilities (full FSArray order)
        self._known_inputs: list = []
        self.clear_iterator_cache()

    def set_circuit(self, circuit):
        previous = self._circuit
        assert not getattr(circuit, "requires_polarization", False), "Circuit must not contain polarized components"
        self._input_state = None
        self._circuit = circuit
        self._umat = circuit.compute_unitary()
        self._u_dev = self._eng().unitary(np.asarray(self._umat, dtype=np.complex128))
        if self._known_inputs and previous is not None and previous.m == circuit.m:
            # same size: keep the deployed inputs, refresh their coefficients with the new unitary
            stale = list(self._known_inputs)
            self._coefs.clear()
            self._probs.clear()
            for st in stale:
                self._compute(st, want_coefs=True)
        else:
            self._reset()

    def set_input_state(self, input_state):
        super().set_input_state(input_state)
        self.preprocess([input_state])

    def clear_mask(self):
        super().clear_mask()

    def preprocess(self, input_list) -> bool:
        new = False
        for st in input_list:
            if st not in self._coefs and st not in self._probs:
                self._compute(st, want_coefs=self._want_coefs_by_default(st))
                new = True
        return new

    # ------------------------------------------------------------------ compute
    def _want_coefs_by_default(self, st) -> bool:
        # coefficients (16 B/state) serve every query; above the cache budget only probabilities are kept and the
        # coefficients are recomputed if an amplitude is requested
        return fsarray.count(st.m, st.n) * 16 <= self._max_cached_bytes // 2

    def _evict(self, need: int):
        def used():
            return sum(t.numel() * t.element_size() for t in list(self._coefs.values()) + list(self._probs.values()))
        for cache in (self._probs, self._coefs):
            for key in list(cache.keys()):
                if used() + need <= self._max_cached_bytes:
                    return
                if key != self._input_state:
                    del cache[key]

    def _compute(self, st, want_coefs: bool):
        eng = self._eng()
        N = fsarray.count(st.m, st.n)
        self._evict(N * (24 if want_coefs else 8))
        occ = [int(x) for x in st]
        probs, psum, coefs = eng.slos_probs(self._u_dev, occ, want_coefs=want_coefs)
        eng.check_status()
        self._probs[st] = probs
        if coefs is not None:
            self._coefs[st] = coefs
        if st not in self._known_inputs:
            self._known_inputs.append(st)

    def _get_coefs(self, st) -> torch.Tensor:
        if st not in self._coefs:
            self._compute(st, want_coefs=True)
        return self._coefs[st]

    def _get_probs(self, st) -> torch.Tensor:
        if st not in self._probs:
            self._compute(st, want_coefs=self._want_coefs_by_default(st))
        return self._probs[st]

    def _mask_indices(self, st):
        """Device int64 ranks (into the full FSArray) of the states the current mask keeps, or None without mask.
        The FSMask predicate runs on the device over the whole rank space (C ABI fock_mask_match)."""
        if self._mask is None:
            return None
        key = (st.m, st.n)
        if key not in self._mask_ranks:
            self._mask_ranks[key] = self._eng().mask_ranks(st.m, st.n, self._mask)
        return self._mask_ranks[key]

    def clear_iterator_cache(self):
        super().clear_iterator_cache()
        self._mask_ranks = {}

    def _get_iterator(self, input_state):
        """Same contract as _abstract_backends.py:148-158 (tuple of output states, cached per photon count).  With a mask
        only the kept ranks are un-ranked, on the device, instead of walking the whole FSArray in Python."""
        n = input_state.n
        if self._mask is None or self._input_state is None or n != self._input_state.n:
            return super()._get_iterator(input_state)
        if n not in self._cache_iterator:
            ranks = self._mask_indices(input_state)
            occ = self._eng().unrank(input_state.m, n, ranks).cpu().numpy()
            self._cache_iterator[n] = tuple(FockState([int(x) for x in row]) for row in occ)
        return self._cache_iterator[n]

    # ------------------------------------------------------------------ reference API (_slos.py:187-223)
    def prob_amplitude(self, output_state) -> complex:
        istate = self._input_state
        if istate.n != output_state.n:
            return complex(0)
        occ = np.array([[int(x) for x in output_state]], dtype=np.uint8)
        idx = int(fsarray.rank_states(istate.m, istate.n, occ)[0])
        assert idx != fsarray.NPOS
        if self._mask is not None:
            assert self._mask.match(output_state), "output state is outside the mask"
        c = complex(self._get_coefs(istate)[idx].item())
        return c * math.sqrt(output_state.prodnfact() / istate.prodnfact())

    def all_prob_tensor(self, input_state=None) -> torch.Tensor:
        """Probabilities of every state of FSArray(m, n) (mask applied if set) as a device float64 tensor."""
        if input_state is not None:
            self.set_input_state(input_state)
        st = self._input_state
        probs = self._get_probs(st)
        idx = self._mask_indices(st)
        if idx is not None:
            probs = probs[idx]
        return probs

    def all_amplitudes_tensor(self, input_state=None) -> torch.Tensor:
        if input_state is not None:
            self.set_input_state(input_state)
        st = self._input_state
        amps = self._eng().slos_amplitudes_from_coefs(st.m, st.n, self._get_coefs(st), prodnfact(st))
        idx = self._mask_indices(st)
        if idx is not None:
            amps = amps[idx]
        return amps

    def coefs_tensor(self, input_state=None) -> torch.Tensor:
        if input_state is not None:
            self.set_input_state(input_state)
        return self._get_coefs(self._input_state)

    def all_prob(self, input_state=None):
        return self.all_prob_tensor(input_state).cpu().tolist()

    def prob_distribution(self):
        probs = self.all_prob_tensor().cpu().tolist()
        bsd = BSDistribution()
        for s, p in zip(self._get_iterator(self._input_state), probs):
            bsd.add(s, p)
        return bsd

    def prob_iterator_tensors(self, min_p: float = MIN_P):
        """(ranks, probabilities) of the states with p > min_p, thresholded and compacted on the device."""
        st = self._input_state
        probs = self._get_probs(st)
        keep = probs > min_p
        idx = self._mask_indices(st)
        if idx is not None:
            allowed = torch.zeros_like(keep)
            allowed[idx] = True
            keep &= allowed
        ranks = torch.nonzero(keep).view(-1)
        return ranks, probs[ranks]

    def prob_iterator(self, min_p: float = MIN_P):
        st = self._input_state
        ranks, probs = self.prob_iterator_tensors(min_p)
        occ = self._eng().unrank(st.m, st.n, ranks).cpu().numpy()
        states = [FockState([int(x) for x in row]) for row in occ]
        return _StateProbView(states, probs.cpu().tolist())

    def evolve(self):
        st = self._input_state
        amps = self.all_amplitudes_tensor().cpu().numpy()
        res = StateVector()
        for s, a in zip(self._get_iterator(st), amps):
            res += s * complex(a)
        return res
