"""Mirrors of Perceval's backend ABCs for installs where Perceval itself cannot be imported.

Same names, attributes, argument meaning and error behaviour as reference perceval/backends/_abstract_backends.py:39-208
(ABackend :39-70, ASamplingBackend :73-81, AStrongSimulationBackend :93-208), written against the local state types.
The generic per-output loops of the reference (all_prob / prob_distribution / prob_iterator / evolve) are kept as the
default behaviour so a subclass only has to provide ``prob_amplitude``; the device backends override them with
batched kernels.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

from ..masks import FockMask
from ..states import BSDistribution, FockState, StateVector


class ABackend(ABC):
    def __init__(self):
        self._circuit = None
        self._umat = None
        self._input_state = None

    def set_circuit(self, circuit):
        # _abstract_backends.py:45-54
        if getattr(circuit, "requires_polarization", False):
            raise RuntimeError("Circuit must not contain polarized components")
        self._input_state = None
        self._circuit = circuit
        self._umat = circuit.compute_unitary()

    def set_input_state(self, input_state):
        self._check_state(input_state)
        self._input_state = input_state

    def _check_state(self, state):
        # _abstract_backends.py:63-65 (pinned by tests/backends/test_backends.py:106-113)
        assert self._circuit is not None, 'Circuit must be set before the input state'
        assert self._circuit.m == state.m, f'Circuit({self._circuit.m}) and state({state.m}) size mismatch'

    @property
    @abstractmethod
    def name(self) -> str:
        """Returns the back-end name as a string"""


class ASamplingBackend(ABackend):
    @abstractmethod
    def sample(self):
        """Request one sample from the circuit given an input state"""

    @abstractmethod
    def samples(self, count: int):
        """Request samples from the circuit given an input state"""


class _StateProbIterator:
    """Re-iterable (state, probability) view (simulators/_simulator_utils.py:148-171 walks it twice)."""

    def __init__(self, states, probs):
        self.states = states
        self.probs = probs

    def __iter__(self):
        return zip(self.states, self.probs)


class AStrongSimulationBackend(ABackend):
    def __init__(self):
        super().__init__()
        self._mask_n = None
        self._cache_iterator: dict = {}
        self._masks_str = None
        self._mask = None
        self._no_limit_modes = None

    # ---- masks (_abstract_backends.py:103-146)
    def set_mask(self, masks, n=None, at_least_modes=None):
        self.clear_mask()
        if isinstance(masks, str):
            masks = [masks]
        width = len(masks[0])
        for msk in masks:
            assert len(msk.replace("*", " ")) == width, "Inconsistent mask lengths"
        self._masks_str = masks
        self._mask_n = n
        self._no_limit_modes = at_least_modes
        self._init_mask()

    def _init_mask(self):
        if self._masks_str is not None and self._input_state is not None:
            st = self._input_state
            assert len(self._masks_str[0]) == st.m, "Mask and input state lengths have to be the same"
            self._mask = FockMask(st.m, self._mask_n or st.n, self._masks_str, self._no_limit_modes or None)

    def clear_mask(self):
        self._masks_str = None
        self._mask = None
        self._mask_n = None
        self.clear_iterator_cache()

    def set_input_state(self, input_state):
        super().set_input_state(input_state)
        self._init_mask()

    def _get_iterator(self, input_state):
        n = input_state.n
        if n not in self._cache_iterator:
            self._cache_iterator[n] = tuple(self._enumerate_states(input_state.m, n))
        return self._cache_iterator[n]

    def _enumerate_states(self, m, n):
        from ..fsarray import iterate_states
        for s in iterate_states(m, n):
            if self._mask is None or self._mask.match(s):
                yield FockState(s)

    def clear_iterator_cache(self):
        self._cache_iterator = {}

    def set_circuit(self, circuit):
        if self._circuit and circuit.m != self._circuit.m:
            self.clear_iterator_cache()
        super().set_circuit(circuit)

    @abstractmethod
    def prob_amplitude(self, output_state) -> complex:
        pass

    def probability(self, output_state) -> float:
        return abs(self.prob_amplitude(output_state)) ** 2

    def all_prob(self, input_state=None) -> list:
        if input_state is not None:
            self.set_input_state(input_state)
        return [self.probability(s) for s in self._get_iterator(self._input_state)]

    def prob_distribution(self):
        bsd = BSDistribution()
        for s in self._get_iterator(self._input_state):
            bsd.add(s, self.probability(s))
        return bsd

    def prob_iterator(self, min_p: float = 1e-16):
        probs = self.all_prob(self._input_state)
        states = [s for i, s in enumerate(self._get_iterator(self._input_state)) if probs[i] > min_p]
        return _StateProbIterator(states, [p for p in probs if p > min_p])

    def evolve(self):
        res = StateVector()
        for s in self._get_iterator(self._input_state):
            res += s * self.prob_amplitude(s)
        return res
