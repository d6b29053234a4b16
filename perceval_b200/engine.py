"""Torch-facing wrapper of the C ABI: owns device tensors, passes raw pointers + the current CUDA stream.

PyTorch is plumbing only (device memory, streams, DLPack); every number is produced by the hand-written sm_100a
kernels in libfock_b200.so.  All entry points raise if CUDA / the library is unavailable -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FockError, check


def _state_u8(state) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(list(state), dtype=np.uint8))


def prodnfact(state) -> float:
    p = 1.0
    for x in state:
        for v in range(2, int(x) + 1):
            p *= v
    return p


class FockEngine:
    """One engine per CUDA device."""

    _engines: dict = {}

    @classmethod
    def get(cls, device=None) -> "FockEngine":
        idx = cls._device_index(device)
        eng = cls._engines.get(idx)
        if eng is None:
            eng = cls(idx)
            cls._engines[idx] = eng
        return eng

    @staticmethod
    def _device_index(device) -> int:
        if not torch.cuda.is_available():
            raise FockError("perceval_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if device is None:
            return torch.cuda.current_device()
        if isinstance(device, int):
            return device
        d = torch.device(device)
        return d.index if d.index is not None else torch.cuda.current_device()

    def __init__(self, index: int):
        self.index = index
        self.device = torch.device("cuda", index)
        self.lib = _lib.load()
        self.ctx = _lib.context(index)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self) -> int:
        return int(self.lib.fock_launch_count(self.ctx))

    def profile_events(self, begin: torch.cuda.Event | None, end: torch.cuda.Event | None):
        """CUDA events recorded around every probability-layer launch of this engine (C ABI fock_profile_events) until
        called again with (None, None).  The events must have been recorded once (torch creates the handle lazily)."""
        if begin is None:
            check(self.lib.fock_profile_events(self.ctx, None, None), "fock_profile_events")
            return
        for ev in (begin, end):
            if not ev.cuda_event:
                ev.record(torch.cuda.current_stream(self.device))
        check(self.lib.fock_profile_events(self.ctx, C.c_void_p(begin.cuda_event), C.c_void_p(end.cuda_event)), "fock_profile_events")

    def check_status(self):
        check(self.lib.fock_check_status(self.ctx, self._stream()), "fock_check_status")

    def unitary(self, u) -> torch.Tensor:
        """m x m complex128 device tensor from numpy / torch / anything exposing __dlpack__ or __array__."""
        if isinstance(u, torch.Tensor):
            t = u
        elif hasattr(u, "__dlpack__") and not isinstance(u, np.ndarray):
            t = torch.from_dlpack(u)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(u, dtype=np.complex128)))
        t = t.to(device=self.device, dtype=torch.complex128).contiguous()
        if t.dim() != 2 or t.shape[0] != t.shape[1]:
            raise ValueError("unitary must be a square matrix")
        return t

    # ------------------------------------------------------------------ FSArray
    def count(self, m: int, n: int) -> int:
        return _lib.count(m, n)

    def rank(self, m: int, n: int, states: torch.Tensor) -> torch.Tensor:
        states = states.to(device=self.device, dtype=torch.uint8).contiguous().view(-1, m)
        out = torch.empty(states.shape[0], dtype=torch.int64, device=self.device)
        check(self.lib.fock_rank(self.ctx, m, n, states.data_ptr(), states.shape[0], out.data_ptr(), self._stream()), "fock_rank")
        return out

    def unrank(self, m: int, n: int, ranks: torch.Tensor) -> torch.Tensor:
        ranks = ranks.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
        out = torch.empty((ranks.shape[0], m), dtype=torch.uint8, device=self.device)
        check(self.lib.fock_unrank(self.ctx, m, n, ranks.data_ptr(), ranks.shape[0], out.data_ptr(), self._stream()), "fock_unrank")
        return out

    def enumerate(self, m: int, n: int, begin: int = 0, end: int | None = None) -> torch.Tensor:
        end = self.count(m, n) if end is None else end
        out = torch.empty((end - begin, m), dtype=torch.uint8, device=self.device)
        check(self.lib.fock_enumerate(self.ctx, m, n, begin, end, out.data_ptr(), self._stream()), "fock_enumerate")
        return out

    # ------------------------------------------------------------------ FSMask
    def mask_flags(self, m: int, n: int, mask, begin: int = 0, end: int | None = None, allow_missing=False,
                   budget: int | None = None) -> torch.Tensor:
        """uint8 device tensor: 1 where state #(begin+i) of FSArray(m, n) matches ``mask`` (a masks.FockMask).
        ``allow_missing``: partial match of intermediate layers; with ``budget`` = b only the partial matches that b more
        photons can still complete (C ABI fock_mask_match, allow_missing = 2 + b)."""
        end = self.count(m, n) if end is None else end
        conds = mask.conds_array()
        flags = torch.empty(end - begin, dtype=torch.uint8, device=self.device)
        am = (2 + int(budget)) if budget is not None else (1 if allow_missing else 0)
        check(self.lib.fock_mask_match(self.ctx, m, n, conds.ctypes.data_as(C.c_void_p), conds.shape[0], mask.at_least_bits(),
                                       am, begin, end, flags.data_ptr(), self._stream()), "fock_mask_match")
        return flags

    def mask_ranks(self, m: int, n: int, mask, budget: int | None = None, chunk: int = 1 << 28) -> torch.Tensor:
        """Ranks (int64, ascending = FSArray order) of the states of FSArray(m, n) the mask keeps (exact match), or with
        ``budget`` = b the states b more photons can still turn into a match (pruned SLOS layers).  The rank space is
        scanned in chunks so that the temporary flags never exceed ``chunk`` bytes."""
        N = self.count(m, n)
        out = []
        for lo in range(0, N, chunk):
            hi = min(N, lo + chunk)
            idx = torch.nonzero(self.mask_flags(m, n, mask, lo, hi, budget=budget)).view(-1)
            out.append(idx + lo if lo else idx)
        return torch.cat(out) if len(out) != 1 else out[0]

    # ------------------------------------------------------------------ SLOS
    def slos_order(self, state) -> list:
        s = _state_u8(state)
        n = int(s.sum())
        out = np.zeros(max(n, 1), dtype=np.int32)
        check(self.lib.slos_order(len(s), s.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)), "slos_order")
        return [int(x) for x in out[:n]]

    def slos_layer(self, m: int, k: int, U: torch.Tensor, mk: int, parent: torch.Tensor, child: torch.Tensor | None = None,
                   parent_begin: int = 0, child_begin: int = 0, child_end: int | None = None) -> torch.Tensor:
        """One layer (C ABI slos_layer).  parent holds ranks [parent_begin, parent_begin+len)."""
        child_end = self.count(m, k) if child_end is None else child_end
        if child is None:
            child = torch.empty(child_end - child_begin, dtype=torch.complex128, device=self.device)
        assert child.numel() >= child_end - child_begin
        check(self.lib.slos_layer(self.ctx, m, k, U.data_ptr(), mk, parent.data_ptr(), parent_begin,
                                  parent_begin + parent.numel(), child.data_ptr(), child_begin, child_end, self._stream()),
              "slos_layer")
        return child

    def slos_layer_probs(self, m: int, k: int, U: torch.Tensor, mk: int, parent: torch.Tensor, in_prodnfact: float,
                         probs: torch.Tensor | None = None, coefs: torch.Tensor | None = None, psum: torch.Tensor | None = None,
                         parent_begin: int = 0, child_begin: int = 0, child_end: int | None = None) -> torch.Tensor:
        child_end = self.count(m, k) if child_end is None else child_end
        if probs is None:
            probs = torch.empty(child_end - child_begin, dtype=torch.float64, device=self.device)
        check(self.lib.slos_layer_probs(self.ctx, m, k, U.data_ptr(), mk, parent.data_ptr(), parent_begin,
                                        parent_begin + parent.numel(), coefs.data_ptr() if coefs is not None else None,
                                        probs.data_ptr(), psum.data_ptr() if psum is not None else None, float(in_prodnfact),
                                        child_begin, child_end, self._stream()), "slos_layer_probs")
        return probs

    def slos_layer_seg(self, m: int, k: int, U: torch.Tensor, mk: int, parent: torch.Tensor, parent_segs, child: torch.Tensor,
                       child_begin: int, child_end: int) -> torch.Tensor:
        """One layer from a SEGMENTED parent (C ABI slos_layer_seg): ``parent`` holds the ranks of ``parent_segs`` (one or
        two (b, e) ranges, ascending) packed back to back; ``child`` receives ranks [child_begin, child_end)."""
        from .partition import seg4, segments_len
        assert parent.numel() >= segments_len(parent_segs) and child.numel() >= child_end - child_begin
        seg = np.array(seg4(parent_segs), dtype=np.uint64)
        check(self.lib.slos_layer_seg(self.ctx, m, k, U.data_ptr(), mk, parent.data_ptr(), seg.ctypes.data_as(C.c_void_p),
                                      child.data_ptr(), child_begin, child_end, self._stream()), "slos_layer_seg")
        return child

    def slos_layer_probs_seg(self, m: int, k: int, U: torch.Tensor, mk: int, parent: torch.Tensor, parent_segs, in_prodnfact: float,
                             probs: torch.Tensor, psum: torch.Tensor | None, child_begin: int, child_end: int) -> torch.Tensor:
        from .partition import seg4, segments_len
        assert parent.numel() >= segments_len(parent_segs) and probs.numel() >= child_end - child_begin
        seg = np.array(seg4(parent_segs), dtype=np.uint64)
        check(self.lib.slos_layer_probs_seg(self.ctx, m, k, U.data_ptr(), mk, parent.data_ptr(), seg.ctypes.data_as(C.c_void_p), None,
                                            probs.data_ptr(), psum.data_ptr() if psum is not None else None, float(in_prodnfact),
                                            child_begin, child_end, self._stream()), "slos_layer_probs_seg")
        return probs

    def slos_layer_slab(self, m: int, k: int, p: int, U: torch.Tensor, mk: int, parent: torch.Tensor, rho_ranges, parent_off, child_off,
                        child: torch.Tensor | None = None, probs: torch.Tensor | None = None, psum: torch.Tensor | None = None,
                        in_prodnfact: float = 1.0):
        """One layer in slab-major layout (C ABI slos_layer_slab; perceval_b200/slab.py): ``rho_ranges`` = [(lo, hi)] per prefix
        weight 0..k, ``parent_off`` / ``child_off`` = element offsets of the parent / child slabs inside ``parent`` and
        ``child`` / ``probs`` (Python ints, taken modulo 2^64)."""
        assert len(rho_ranges) == k + 1 and len(parent_off) >= k and len(child_off) >= k + 1
        mask = (1 << 64) - 1
        rr = np.array([x & mask for ab in rho_ranges for x in ab], dtype=np.uint64)
        po = np.array([x & mask for x in parent_off[:k]] or [0], dtype=np.uint64)
        co = np.array([x & mask for x in child_off[:k + 1]], dtype=np.uint64)
        check(self.lib.slos_layer_slab(self.ctx, m, k, p, U.data_ptr(), mk, parent.data_ptr(), child.data_ptr() if child is not None else None,
                                       probs.data_ptr() if probs is not None else None, psum.data_ptr() if psum is not None else None,
                                       float(in_prodnfact), rr.ctypes.data_as(C.c_void_p), po.ctypes.data_as(C.c_void_p),
                                       co.ctypes.data_as(C.c_void_p), self._stream()), "slos_layer_slab")

    def slos_layer_masked(self, m: int, k: int, U: torch.Tensor, mk: int, parent_ranks: torch.Tensor, parent: torch.Tensor,
                          child_ranks: torch.Tensor, in_prodnfact: float = 1.0, want_coefs: bool = True, want_probs: bool = False,
                          want_amps: bool = False, psum: torch.Tensor | None = None):
        """One layer over a pruned rank space (C ABI slos_layer_masked): ``parent`` is packed in the order of the ascending
        int64 ``parent_ranks``; returns (coefs|None, probs|None, amps|None) packed in the order of ``child_ranks``."""
        nc = child_ranks.numel()
        coefs = torch.empty(nc, dtype=torch.complex128, device=self.device) if want_coefs else None
        probs = torch.empty(nc, dtype=torch.float64, device=self.device) if want_probs else None
        amps = torch.empty(nc, dtype=torch.complex128, device=self.device) if want_amps else None
        assert parent.numel() == parent_ranks.numel() and parent_ranks.dtype == torch.int64 and child_ranks.dtype == torch.int64
        check(self.lib.slos_layer_masked(self.ctx, m, k, U.data_ptr(), mk, parent_ranks.data_ptr(), parent_ranks.numel(), parent.data_ptr(),
                                         child_ranks.data_ptr(), nc, coefs.data_ptr() if want_coefs else None,
                                         probs.data_ptr() if want_probs else None, amps.data_ptr() if want_amps else None,
                                         psum.data_ptr() if psum is not None else None, float(in_prodnfact), self._stream()),
              "slos_layer_masked")
        return coefs, probs, amps

    def slos_probs_windowed(self, U: torch.Tensor, in_state, child_begin: int, child_end: int, probs: torch.Tensor | None = None,
                            psum: torch.Tensor | None = None, plan=None, buffers=None, last_events: list | None = None):
        """Probabilities of the ranks [child_begin, child_end) of the output layer, keeping resident -- and computing --
        only the parents of every layer that this range needs (partition.plan_chain): the recompute-window partition
        of a multi-GPU run.  Returns (probs, psum, plan)."""
        from . import partition as P
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        assert n >= 1
        order = self.slos_order(s)
        if plan is None:
            plan = P.plan_chain(m, n, child_begin, child_end)
        if psum is None:
            psum = torch.zeros(1, dtype=torch.float64, device=self.device)
        if child_end <= child_begin:
            return (probs if probs is not None else torch.empty(0, dtype=torch.float64, device=self.device)), psum, plan
        # two packed ping-pong buffers: layers n-1, n-3, ... in A; n-2, n-4, ... in B
        need_a = max([P.segments_len(plan[k]) for k in range(n - 1, -1, -2)] + [1])
        need_b = max([P.segments_len(plan[k]) for k in range(n - 2, -1, -2)] + [1])
        if buffers is None:
            buffers = (torch.empty(need_a, dtype=torch.complex128, device=self.device),
                       torch.empty(need_b, dtype=torch.complex128, device=self.device))
            own_buffers = True
        else:
            own_buffers = False
        buf_a, buf_b = buffers
        del buffers
        assert buf_a.numel() >= need_a and buf_b.numel() >= need_b
        cur = buf_a if ((n - 1) - 0) % 2 == 0 else buf_b   # layer k lives in A when (n-1-k) is even
        cur[:1] = 1.0                                # layer 0 = the vacuum coefficient
        parent, parent_segs = cur, [(0, 1)]
        child_buf = None
        for k in range(1, n):
            child_buf = buf_a if ((n - 1) - k) % 2 == 0 else buf_b
            off = 0
            for b, e in plan[k]:
                self.slos_layer_seg(m, k, U, order[k - 1], parent, parent_segs, child_buf[off:off + (e - b)], b, e)
                off += e - b
            parent, parent_segs = child_buf, plan[k]
        if probs is None:
            if own_buffers:   # layer n-1 sits in A: drop B before the probabilities are allocated (peak = max(A+B, A+probs))
                del buf_b, child_buf, cur
                torch.cuda.synchronize(self.device)
                torch.cuda.empty_cache()
            probs = torch.empty(child_end - child_begin, dtype=torch.float64, device=self.device)
        if last_events is not None:   # CUDA events around the dominant launch (bench.py roofline)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self.slos_layer_probs_seg(m, n, U, order[n - 1], parent, parent_segs, prodnfact(s), probs, psum, child_begin, child_end)
        if last_events is not None:
            ev[1].record()
            last_events.append((ev[0], ev[1], 16 * P.segments_len(parent_segs) + 8 * (child_end - child_begin)))
        return probs, psum, plan

    def slos_coefs(self, U: torch.Tensor, in_state) -> torch.Tensor:
        """Un-normalised SLOS coefficients of the last layer (what _Path.coefs holds, _slos.py:44)."""
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        order = self.slos_order(s)
        cur = torch.ones(1, dtype=torch.complex128, device=self.device)
        for k in range(1, n + 1):
            cur = self.slos_layer(m, k, U, order[k - 1], cur)
        return cur

    def slos_probs(self, U: torch.Tensor, in_state, want_coefs: bool = False, workspaces=None):
        """Full output distribution in FSArray order.  Returns (probs, sum, coefs|None).

        Runs the whole chain through the C ABI ``slos_prob_distribution`` with ping-pong workspaces sized for
        layers n-1 and n-2 (only two layers are ever live)."""
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        N = self.count(m, n)
        if workspaces is None:
            wa = torch.empty(self.count(m, n - 1) if n >= 1 else 1, dtype=torch.complex128, device=self.device)
            wb = torch.empty(self.count(m, n - 2) if n >= 2 else 1, dtype=torch.complex128, device=self.device)
        else:
            wa, wb = workspaces
        probs = torch.empty(N, dtype=torch.float64, device=self.device)
        psum = torch.zeros(1, dtype=torch.float64, device=self.device)
        coefs = torch.empty(N, dtype=torch.complex128, device=self.device) if want_coefs else None
        check(self.lib.slos_prob_distribution(self.ctx, m, U.data_ptr(), s.ctypes.data_as(C.c_void_p), wa.data_ptr(), wb.data_ptr(),
                                              coefs.data_ptr() if coefs is not None else None, probs.data_ptr(), psum.data_ptr(),
                                              self._stream()), "slos_prob_distribution")
        return probs, psum, coefs

    def slos_probs_to_host(self, U: torch.Tensor, in_state, out: torch.Tensor, pieces: int = 8, child_begin: int = 0,
                           child_end: int | None = None, workspaces=None, probs: torch.Tensor | None = None,
                           parent: torch.Tensor | None = None):
        """Output distribution (ranks [child_begin, child_end) of FSArray order) straight into the HOST tensor ``out``
        (pinned memory for an asynchronous copy).  The last layer is cut into ``pieces`` rank ranges; the device->host
        copy of piece i runs on a side stream while piece i+1 is computed, so the PCIe transfer (6.7 GB at 12 photons /
        24 modes) hides the last-layer kernel instead of following it.  ``parent`` = layer n-1 if the caller already
        holds it (sharded chains).  Returns sum(p) over the range as a 1-element device tensor."""
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        assert n >= 1
        N = self.count(m, n)
        child_end = N if child_end is None else child_end
        total = child_end - child_begin
        assert out.numel() >= total and out.dtype == torch.float64 and not out.is_cuda
        order = self.slos_order(s)
        if parent is None:
            if workspaces is None:
                workspaces = (torch.empty(self.count(m, n - 1), dtype=torch.complex128, device=self.device),
                              torch.empty(max(self.count(m, n - 2), 1) if n >= 2 else 1, dtype=torch.complex128, device=self.device))
            wa, wb = workspaces
            parent = torch.ones(1, dtype=torch.complex128, device=self.device)
            for k in range(1, n):
                nc = self.count(m, k)
                buf = wa if (n - 1 - k) % 2 == 0 else wb
                parent = self.slos_layer(m, k, U, order[k - 1], parent, child=buf[:nc])[:nc]
        if probs is None:
            probs = torch.empty(total, dtype=torch.float64, device=self.device)
        psum = torch.zeros(1, dtype=torch.float64, device=self.device)
        inf = prodnfact(s)
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
        side = self._copy_stream
        pieces = max(1, min(int(pieces), total))
        for i in range(pieces):
            b = child_begin + total * i // pieces
            e = child_begin + total * (i + 1) // pieces
            if e == b:
                continue
            dst = probs[b - child_begin:e - child_begin]
            self.slos_layer_probs(m, n, U, order[n - 1], parent, inf, probs=dst, psum=psum, child_begin=b, child_end=e)
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)
            with torch.cuda.stream(side):
                out[b - child_begin:e - child_begin].copy_(dst, non_blocking=True)
        main.wait_stream(side)
        return psum

    def slos_probs_from_coefs(self, m: int, n: int, coefs: torch.Tensor, in_prodnfact: float):
        probs = torch.empty(coefs.numel(), dtype=torch.float64, device=self.device)
        psum = torch.zeros(1, dtype=torch.float64, device=self.device)
        check(self.lib.slos_probs_epilogue(self.ctx, m, n, coefs.data_ptr(), float(in_prodnfact), probs.data_ptr(), psum.data_ptr(),
                                           0, coefs.numel(), self._stream()), "slos_probs_epilogue")
        return probs, psum

    def slos_amplitudes_from_coefs(self, m: int, n: int, coefs: torch.Tensor, in_prodnfact: float) -> torch.Tensor:
        amps = torch.empty(coefs.numel(), dtype=torch.complex128, device=self.device)
        check(self.lib.slos_amplitudes_epilogue(self.ctx, m, n, coefs.data_ptr(), float(in_prodnfact), amps.data_ptr(), 0,
                                                coefs.numel(), self._stream()), "slos_amplitudes_epilogue")
        return amps

    def slos_probs_host(self, u: np.ndarray, in_state) -> tuple[np.ndarray, float]:
        """End-to-end call with HOST buffers through the C ABI (U in, probabilities out)."""
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        u = np.ascontiguousarray(np.asarray(u, dtype=np.complex128))
        probs = np.empty(self.count(m, n), dtype=np.float64)
        psum = np.zeros(1, dtype=np.float64)
        check(self.lib.slos_prob_distribution_host(self.ctx, m, u.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                                   probs.ctypes.data_as(C.c_void_p), psum.ctypes.data_as(C.c_void_p)),
              "slos_prob_distribution_host")
        return probs, float(psum[0])

    # ------------------------------------------------------------------ permanents / Naive
    def permanents(self, mats: torch.Tensor, gray_begin: int = 0, gray_end: int = 0) -> torch.Tensor:
        """Batched permanents of (B, n, n) complex128 matrices (Glynn); optional Gray-code sub-range."""
        mats = mats.to(device=self.device, dtype=torch.complex128).contiguous()
        if mats.dim() == 2:
            mats = mats.unsqueeze(0)
        B, n = mats.shape[0], mats.shape[1]
        assert mats.shape[2] == n
        out = torch.empty(B, dtype=torch.complex128, device=self.device)
        check(self.lib.glynn_permanent_batch(self.ctx, n, mats.data_ptr(), B, out.data_ptr(), gray_begin, gray_end, self._stream()),
              "glynn_permanent_batch")
        return out

    def permanents_host(self, mats: np.ndarray) -> np.ndarray:
        mats = np.ascontiguousarray(np.asarray(mats, dtype=np.complex128))
        if mats.ndim == 2:
            mats = mats[None]
        out = np.empty(mats.shape[0], dtype=np.complex128)
        check(self.lib.glynn_permanent_batch_host(self.ctx, mats.shape[1], mats.ctypes.data_as(C.c_void_p), mats.shape[0],
                                                  out.ctypes.data_as(C.c_void_p)), "glynn_permanent_batch_host")
        return out

    def naive_amplitudes(self, U: torch.Tensor, in_state, out_ranks: torch.Tensor | None = None,
                         out_states: torch.Tensor | None = None) -> torch.Tensor:
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        if out_states is not None:
            st = out_states.to(device=self.device, dtype=torch.uint8).contiguous().view(-1, m)
            amps = torch.empty(st.shape[0], dtype=torch.complex128, device=self.device)
            check(self.lib.naive_amplitudes_states(self.ctx, m, n, U.data_ptr(), s.ctypes.data_as(C.c_void_p), st.data_ptr(),
                                                   st.shape[0], amps.data_ptr(), self._stream()), "naive_amplitudes_states")
            return amps
        rk = out_ranks.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
        amps = torch.empty(rk.shape[0], dtype=torch.complex128, device=self.device)
        check(self.lib.naive_amplitudes(self.ctx, m, n, U.data_ptr(), s.ctypes.data_as(C.c_void_p), rk.data_ptr(), rk.shape[0],
                                        amps.data_ptr(), self._stream()), "naive_amplitudes")
        return amps

    # ------------------------------------------------------------------ Clifford & Clifford
    def cc2017_samples(self, U: torch.Tensor, in_state, count: int, seed: int = 0, offset: int = 0,
                       out: torch.Tensor | None = None) -> torch.Tensor:
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        if out is None:
            out = torch.empty((count, m), dtype=torch.uint8, device=self.device)
        check(self.lib.cc2017_samples(self.ctx, m, n, U.data_ptr(), s.ctypes.data_as(C.c_void_p), count, seed & ((1 << 64) - 1),
                                      offset, out.data_ptr(), self._stream()), "cc2017_samples")
        return out

    def cc2017_samples_host(self, u: np.ndarray, in_state, count: int, seed: int = 0, offset: int = 0) -> np.ndarray:
        s = _state_u8(in_state)
        m, n = len(s), int(s.sum())
        u = np.ascontiguousarray(np.asarray(u, dtype=np.complex128))
        out = np.empty((count, m), dtype=np.uint8)
        check(self.lib.cc2017_samples_host(self.ctx, m, n, u.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), count,
                                           seed & ((1 << 64) - 1), offset, out.ctypes.data_as(C.c_void_p)), "cc2017_samples_host")
        return out

    # ------------------------------------------------------------------ measurement
    def measure_peak(self, kind: int) -> float:
        v = C.c_double(0)
        check(self.lib.fock_measure_peak(self.ctx, kind, C.byref(v)), "fock_measure_peak")
        return float(v.value)
