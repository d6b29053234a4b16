"""SLOS_B200 -- device SLOS backend behind Perceval's AStrongSimulationBackend surface.

Mirrors reference perceval/backends/_slos.py:105-223 (SLOSBackend): same constructor keyword (``mask``), same method
names, argument meaning and error behaviour; every number comes from libfock_b200.so (slos_layer / slos_layer_probs /
slos_layer_masked / epilogues).  What is built the same way as the reference:

  * ``preprocess(list)`` deploys ONE compute tree for all new input states (``_PathNode`` = the reference's ``_Path``,
    _slos.py:36-86, same greedy rule: the mode most inputs still need comes first), so inputs that share photons share the
    layers of their common prefix -- a noisy source's dozens of input states (simulators/simulator.py:458-471) cost one
    chain plus their differing tails, not one chain each;
  * a same-size circuit change keeps the deployed inputs and refreshes their coefficients (_slos.py:136-139, pinned by
    tests/backends/test_backends.py:252-277) -- lazily, at the next query, instead of eagerly for every known input;
  * with a mask every layer lives on the pruned rank space the reference builds with xq.FSArray(m, k, mask)
    (_slos.py:156-166): only states that can still grow into an accepted output are stored or computed.

Differences that are invisible through the ABC contract:
  * a node's layer is freed as soon as its last child is computed (the reference keeps all n, _slos.py:44); results are
    cached per input state on the device (the reference's ``_state_mapping``) within ``max_cached_bytes``;
  * large distributions are kept as float64 probabilities only (fused last-layer epilogue); coefficients are recomputed if
    an amplitude is asked for;
  * ``device_ids=[...]`` shards the output layer of a chain over several GPUs of this process (per-device streams, no
    torchrun): ``Processor("SLOS_B200")`` stays a plain constructor call;
  * tensor-returning variants (``all_prob_tensor``, ``all_prob_into``, ``all_amplitudes_tensor``, ``prob_iterator_tensors``)
    avoid the one-Python-object-per-state result types that cannot carry 8e8 states.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import fsarray
from .._compat import MIN_P, AStrongSimulationBackend, BSDistribution, FockState, StateVector
from ..engine import FockEngine, prodnfact
from ..masks import FockMask


class _StateProbView:
    """Re-iterable (state, probability) pairs (reference _abstract_backends.py:83-90)."""

    def __init__(self, states, probs):
        self.states = states
        self.probs = probs

    def __iter__(self):
        return zip(self.states, self.probs)

    def __len__(self):
        return len(self.probs)


class _PathNode:
    """Node of the minimal computing graph covering a set of input states (reference _slos.py:36-86, ``_Path``): the node
    at depth k holds the layer-k coefficients of a partial input; ``children[mode]`` adds one photon in ``mode``."""

    __slots__ = ("depth", "children", "states", "nmax")

    def __init__(self, depth: int, states, targets):
        self.depth = depth
        self.children: dict = {}
        self.states: list = []
        self.nmax = depth
        rem_t, rem_s = [], []
        for t, s in zip(targets, states):
            if sum(t) == 0:
                self.states.append(s)            # _slos.py:54-55: this input ends here
            else:
                rem_t.append(t)
                rem_s.append(s)
        # _decompose, _slos.py:61-86: repeatedly split off the inputs that still hold a photon in the most requested mode
        while rem_t:
            counts = [sum(col) for col in zip(*rem_t)]
            mode = counts.index(max(counts))
            cur_t, cur_s, new_t, new_s = [], [], [], []
            for t, s in zip(rem_t, rem_s):
                if t[mode]:
                    t[mode] -= 1
                    cur_t.append(t)
                    cur_s.append(s)
                else:
                    new_t.append(t)
                    new_s.append(s)
            child = _PathNode(depth + 1, cur_s, cur_t)
            self.children[mode] = child
            self.nmax = max(self.nmax, child.nmax)
            rem_t, rem_s = new_t, new_s

    def count_layers(self) -> int:
        return sum(1 + c.count_layers() for c in self.children.values())


class _Result:
    """Cached output of one input state: ``ranks`` = kept ranks (None: the whole FSArray, in order)."""

    __slots__ = ("ranks", "coefs", "probs", "psum", "shards")

    def __init__(self, ranks=None, coefs=None, probs=None, psum=None, shards=None):
        self.ranks, self.coefs, self.probs, self.psum, self.shards = ranks, coefs, probs, psum, shards

    def nbytes(self) -> int:
        tot = 0
        for t in (self.coefs, self.probs):
            if t is not None:
                tot += t.numel() * t.element_size()
        for t in self.shards or ():
            tot += t.numel() * t.element_size()
        return tot


class SLOSB200Backend(AStrongSimulationBackend):
    def __init__(self, mask=None, device=None, device_ids=None, max_cached_bytes: int = 16 << 30, use_symbolic: bool = False,
                 lazy_above: int = 1 << 26):
        """``device_ids``: GPUs of this process over which the output layer of a chain is sharded (default: one device).
        ``lazy_above``: an un-masked single input with more output states than this is computed at its first query
        instead of inside set_input_state, so that ``all_prob_into`` can pipeline the last layer with the device->host copy."""
        super().__init__()
        if use_symbolic:
            raise NotImplementedError("SLOS_B200 is numeric (complex128) only; use the reference SLOS backend for sympy")
        if device_ids is not None:
            device_ids = [int(d) for d in device_ids]
            assert len(device_ids) >= 1 and len(set(device_ids)) == len(device_ids), "device_ids must be distinct"
            device = device_ids[0] if device is None else device
        self._device = device
        self._device_ids = device_ids if device_ids and len(device_ids) > 1 else None
        self._engine: FockEngine | None = None
        self._u_dev = None
        self._u_peers: dict = {}
        self._max_cached_bytes = max_cached_bytes
        self._lazy_above = lazy_above
        self._dev_mask: FockMask | None = None
        self.stats = {"layers_computed": 0, "trees": 0}
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
        self._results: dict = {}      # input state -> _Result (reference: _state_mapping)
        self._stale: set = set()      # deployed inputs whose coefficients belong to a previous unitary
        self._known_inputs: list = []
        self._kept: dict = {}         # (k, budget) -> ascending kept ranks of FSArray(m, k) under the current mask
        self.clear_iterator_cache()

    def set_circuit(self, circuit):
        previous = self._circuit
        assert not getattr(circuit, "requires_polarization", False), "Circuit must not contain polarized components"
        self._input_state = None
        self._circuit = circuit
        self._umat = circuit.compute_unitary()
        self._u_dev = self._eng().unitary(self._umat)
        self._u_peers = {}
        if self._known_inputs and previous is not None and previous.m == circuit.m:
            # same size: keep the deployed inputs; their coefficients are refreshed with the new unitary when next queried
            self._stale = {st for st in self._known_inputs if st in self._results}
            self._known_inputs = [st for st in self._known_inputs if st in self._stale]
            self._results.clear()
        else:
            self._reset()

    def set_input_state(self, input_state):
        super().set_input_state(input_state)
        if self._defer(input_state):
            if input_state not in self._known_inputs:
                self._known_inputs.append(input_state)
            return
        self.preprocess([input_state])

    def _init_mask(self):
        # With a real Perceval install the base class builds an exqalibur xq.FSMask (used for .match()); the device path
        # needs the conditions themselves, so a local FockMask is built next to it from the same fields
        # (_abstract_backends.py:130-137).
        super()._init_mask()
        self._dev_mask = None
        if self._masks_str is not None and self._input_state is not None:
            st = self._input_state
            self._dev_mask = FockMask(st.m, self._mask_n or st.n, [msk.replace("*", " ") for msk in self._masks_str],
                                      self._no_limit_modes or None)

    def set_mask(self, masks, n=None, at_least_modes=None):
        super().set_mask(masks, n=n, at_least_modes=at_least_modes)
        self._drop_results()

    def clear_mask(self):
        # _slos.py:147-149: a mask change invalidates every deployed path
        super().clear_mask()
        self._dev_mask = None
        if hasattr(self, "_results"):
            self._drop_results()

    def _drop_results(self):
        self._stale |= set(self._results.keys())
        self._results.clear()
        self._kept = {}
        self.clear_iterator_cache()

    # ------------------------------------------------------------------ planning
    def _want_coefs_by_default(self, st) -> bool:
        # coefficients (16 B/state) serve every query; above the cache budget only probabilities are kept and the
        # coefficients are recomputed if an amplitude is requested
        if self._dev_mask is not None:
            return True
        return fsarray.count(st.m, st.n) * 16 <= self._max_cached_bytes // 2

    def _defer(self, st) -> bool:
        return (self._dev_mask is None and self._masks_str is None and st not in self._results
                and fsarray.count(st.m, st.n) > self._lazy_above and not self._want_coefs_by_default(st))

    def _evict(self, need: int, keep=()):
        def used():
            return sum(r.nbytes() for r in self._results.values())
        for key in list(self._results.keys()):
            if used() + need <= self._max_cached_bytes:
                return
            if key != self._input_state and key not in keep:
                del self._results[key]

    def preprocess(self, input_list) -> bool:
        """Deploys one compute tree for every input of ``input_list`` that has no current result (reference
        _slos.py:170-185) and computes it.  Returns False if nothing was new."""
        new = []
        for st in input_list:
            if st not in self._results and st not in new:
                new.append(st)
        if not new:
            return False
        m = self._circuit.m
        for st in new:
            assert st.m == m, f"Circuit({m}) and state({st.m}) size mismatch"
        self._run_tree(new, {st: self._want_coefs_by_default(st) for st in new})
        return True

    def _kept_ranks(self, m: int, k: int, budget: int | None):
        """Ascending kept ranks of layer k under the current mask: exact matches (budget None) or the states ``budget``
        more photons can still complete (what xq.FSArray(m, k, mask) holds, _slos.py:156-166)."""
        key = (k, budget)
        if key not in self._kept:
            self._kept[key] = self._eng().mask_ranks(m, k, self._dev_mask, budget=budget)
        return self._kept[key]

    # ------------------------------------------------------------------ compute
    def _run_tree(self, states, want_coefs: dict):
        eng = self._eng()
        m = self._circuit.m
        need = 0
        for st in states:
            N = fsarray.count(st.m, st.n)
            need += N * (24 if want_coefs[st] else 8) if self._dev_mask is None else 0
        self._evict(need, keep=states)
        root = _PathNode(0, list(states), [[int(x) for x in st] for st in states])
        self.stats["trees"] += 1
        if (self._dev_mask is None and len(states) == 1 and self._device_ids is None and states[0].n >= 1):
            # a single chain: the whole run is one C-ABI call (slos_prob_distribution) on ping-pong workspaces
            st = states[0]
            probs, psum, coefs = eng.slos_probs(self._u_dev, [int(x) for x in st], want_coefs=want_coefs[st])
            self.stats["layers_computed"] += st.n
            self._store(st, _Result(None, coefs, probs, psum))
        elif self._dev_mask is None and len(states) == 1 and self._device_ids is not None and states[0].n >= 1:
            self._run_chain_sharded(states[0])
        else:
            vac = torch.ones(1, dtype=torch.complex128, device=eng.device)
            ranks0 = torch.zeros(1, dtype=torch.int64, device=eng.device) if self._dev_mask is not None else None
            for st in root.states:                  # the vacuum input (n = 0): one state, coefficient 1
                self._finish(st, 0, ranks0, vac, want_coefs[st])
            self._descend(root, [(ranks0, vac)], m, want_coefs)
        if self._dev_mask is not None:
            eng.check_status()      # synchronises; only the pruned kernels can flag an error (a kept child without its parent)

    def _descend(self, node: _PathNode, holder: list, m: int, want_coefs: dict):
        eng = self._eng()
        children = list(node.children.items())
        for i, (mk, child) in enumerate(children):
            pranks, parent = holder[0]
            k = child.depth
            leaf = not child.children
            self.stats["layers_computed"] += 1
            rep = None
            if self._dev_mask is None:
                st_here = child.states[0] if child.states else None
                if leaf and st_here is not None and not want_coefs[st_here]:
                    psum = torch.zeros(1, dtype=torch.float64, device=eng.device)
                    probs = eng.slos_layer_probs(m, k, self._u_dev, mk, parent, prodnfact(st_here), psum=psum)
                    for st in child.states:
                        self._store(st, _Result(None, None, probs, psum))
                else:
                    coefs = eng.slos_layer(m, k, self._u_dev, mk, parent)
                    for st in child.states:
                        self._finish(st, k, None, coefs, want_coefs[st])
                    rep = (None, coefs)
            else:
                cranks = self._kept_ranks(m, k, max(child.nmax, self._mask_n or 0) - k)
                coefs, _, _ = eng.slos_layer_masked(m, k, self._u_dev, mk, pranks, parent, cranks)
                for st in child.states:
                    self._finish(st, k, cranks, coefs, want_coefs[st])
                rep = (cranks, coefs)
            del pranks, parent
            if i == len(children) - 1:
                holder[0] = None                      # the node's layer is no longer needed: free it before going deeper
            if not leaf:
                self._descend(child, [rep], m, want_coefs)
            del rep

    def _finish(self, st, k: int, ranks, coefs, keep_coefs: bool):
        """Turns the layer an input ends on into its cached result (probabilities always, coefficients if wanted)."""
        eng = self._eng()
        inf = prodnfact(st)
        if ranks is None:
            probs, psum = eng.slos_probs_from_coefs(st.m, k, coefs, inf)
            self._store(st, _Result(None, coefs if keep_coefs else None, probs, psum))
            return
        # pruned layer: restrict to the states this input keeps (the node may carry partial matches for longer inputs)
        exact = self._mask_indices(st)
        if exact.numel() != ranks.numel():
            pos = torch.searchsorted(ranks, exact)
            coefs = coefs[pos]
        probs = self._masked_probs(st, exact, coefs, inf)
        self._store(st, _Result(exact, coefs, probs, probs.sum().view(1)))

    def _masked_probs(self, st, ranks, coefs, inf):
        if ranks.numel() == 0:
            return torch.empty(0, dtype=torch.float64, device=coefs.device)
        occ = self._eng().unrank(st.m, st.n, ranks).to(torch.float64)
        fact = torch.exp(torch.lgamma(occ + 1.0).sum(dim=1)).round()
        return (coefs.real ** 2 + coefs.imag ** 2) * (fact / inf)

    def _store(self, st, res: _Result):
        self._results[st] = res
        self._stale.discard(st)
        if st not in self._known_inputs:
            self._known_inputs.append(st)

    # ---- several GPUs in one process: replicated lower layers, output layer sharded by rank range
    def _peer(self, dev: int):
        if dev not in self._u_peers:
            e = FockEngine.get(dev)
            self._u_peers[dev] = (e, self._u_dev.to(e.device))
        return self._u_peers[dev]

    def _run_chain_sharded(self, st):
        """One chain on ``device_ids``: every device computes layers 1..n-1 on its own stream (replicated: recomputing
        them is cheaper than moving them) and its rank range of the output layer; the shards stay on
        their devices (``all_prob_shards``) until a gathered tensor is asked for."""
        from ..dist import shard_range
        occ = [int(x) for x in st]
        m, n = st.m, st.n
        N = fsarray.count(m, n)
        W = len(self._device_ids)
        shards, sums = [], []
        for r, dev in enumerate(self._device_ids):
            e, U = self._peer(dev)
            b, en = shard_range(N, r, W)
            with torch.cuda.device(e.device):
                order = e.slos_order(occ)
                parent = torch.ones(1, dtype=torch.complex128, device=e.device)
                for k in range(1, n):
                    parent = e.slos_layer(m, k, U, order[k - 1], parent)
                psum = torch.zeros(1, dtype=torch.float64, device=e.device)
                shards.append(e.slos_layer_probs(m, n, U, order[n - 1], parent, prodnfact(occ), psum=psum, child_begin=b, child_end=en))
                sums.append(psum)
                del parent
        self.stats["layers_computed"] += n
        for dev in self._device_ids:
            FockEngine.get(dev).check_status()
        main = self._eng().device
        psum = torch.stack([s_.to(main) for s_ in sums]).sum(dim=0)
        self._store(st, _Result(None, None, None, psum, shards))

    # ------------------------------------------------------------------ cached results
    def _result(self, st, need_coefs: bool = False) -> _Result:
        res = self._results.get(st)
        if res is None or (need_coefs and res.coefs is None):
            self._run_tree([st], {st: True if need_coefs else self._want_coefs_by_default(st)})
            res = self._results[st]
        return res

    def _get_coefs(self, st) -> torch.Tensor:
        return self._result(st, need_coefs=True).coefs

    def _get_probs(self, st) -> torch.Tensor:
        res = self._result(st)
        if res.probs is None and res.shards is not None:
            main = self._eng().device
            res.probs = torch.cat([s_.to(main) for s_ in res.shards])
        return res.probs

    def all_prob_shards(self, input_state=None):
        """[(device, (begin, end), float64 tensor)] -- the output distribution as it lives on ``device_ids`` (one entry on a
        single device)."""
        from ..dist import shard_range
        if input_state is not None:
            self.set_input_state(input_state)
        st = self._input_state
        res = self._result(st)
        if res.shards is None:
            p = self._get_probs(st)
            return [(p.device, (0, p.numel()), p)]
        N = fsarray.count(st.m, st.n)
        return [(s_.device, shard_range(N, r, len(res.shards)), s_) for r, s_ in enumerate(res.shards)]

    def _mask_indices(self, st):
        """Device int64 ranks (into the full FSArray) of the states the current mask keeps, or None without mask."""
        if self._dev_mask is None:
            return None
        # a mask instantiated for more photons than this input holds (set_mask(..., n=...): the input is one part of a
        # separated state) keeps the states the missing photons can still complete
        missing = (self._mask_n or 0) - st.n
        return self._kept_ranks(st.m, st.n, missing if missing > 0 else None)

    def clear_iterator_cache(self):
        super().clear_iterator_cache()

    def _get_iterator(self, input_state):
        """Same contract as _abstract_backends.py:148-158 (tuple of output states, cached per photon count).  With a mask
        only the kept ranks are un-ranked, on the device, instead of walking the whole FSArray in Python."""
        n = input_state.n
        if self._dev_mask is None or self._input_state is None or n != self._input_state.n:
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
        res = self._result(istate, need_coefs=True)
        if res.ranks is not None:
            pos = int(torch.searchsorted(res.ranks, torch.tensor([idx], dtype=torch.int64, device=res.ranks.device)).item())
            assert pos < res.ranks.numel() and int(res.ranks[pos].item()) == idx, "output state is outside the mask"
            idx = pos
        c = complex(res.coefs[idx].item())
        return c * math.sqrt(output_state.prodnfact() / istate.prodnfact())

    def all_prob_tensor(self, input_state=None) -> torch.Tensor:
        """Probabilities of every state of FSArray(m, n) (of the kept states if a mask is set) as a device float64 tensor."""
        if input_state is not None:
            self.set_input_state(input_state)
        return self._get_probs(self._input_state)

    def all_prob_into(self, out: torch.Tensor, input_state=None, pieces: int = 8) -> float:
        """The same distribution written into the HOST float64 tensor ``out`` (pinned memory for asynchronous copies);
        returns sum(p).  An input that has not been computed yet runs its last layer in ``pieces`` rank ranges whose
        device->host copies overlap the next piece's kernel (FockEngine.slos_probs_to_host); a cached result is copied."""
        if input_state is not None:
            self.set_input_state(input_state)
        st = self._input_state
        eng = self._eng()
        assert not out.is_cuda and out.dtype == torch.float64
        if st not in self._results and self._dev_mask is None and self._device_ids is None and st.n >= 1:
            psum = eng.slos_probs_to_host(self._u_dev, [int(x) for x in st], out, pieces=pieces)
            self.stats["layers_computed"] += st.n
            total = float(psum.item())      # synchronises: the host buffer is complete
            eng.check_status()
            return total
        res = self._result(st)
        if res.shards is not None:
            off = 0
            for s_ in res.shards:
                out[off:off + s_.numel()].copy_(s_, non_blocking=True)
                off += s_.numel()
            for dev in self._device_ids:
                torch.cuda.synchronize(dev)
        else:
            out[:res.probs.numel()].copy_(res.probs, non_blocking=True)
            torch.cuda.synchronize(eng.device)
        return float(res.psum.item()) if res.psum is not None else float(out.sum())

    def all_amplitudes_tensor(self, input_state=None) -> torch.Tensor:
        if input_state is not None:
            self.set_input_state(input_state)
        st = self._input_state
        res = self._result(st, need_coefs=True)
        if res.ranks is None:
            return self._eng().slos_amplitudes_from_coefs(st.m, st.n, res.coefs, prodnfact(st))
        occ = self._eng().unrank(st.m, st.n, res.ranks).to(torch.float64)
        fact = torch.exp(torch.lgamma(occ + 1.0).sum(dim=1)).round()
        return res.coefs * torch.sqrt(fact / prodnfact(st))

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
        res = self._result(st)
        probs = self._get_probs(st)
        idx = torch.nonzero(probs > min_p).view(-1)
        ranks = idx if res.ranks is None else res.ranks[idx]
        return ranks, probs[idx]

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
