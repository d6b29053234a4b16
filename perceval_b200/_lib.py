"""ctypes binding of libfock_b200.so (the C ABI declared in include/fock_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that is impossible, or if a
compute entry point is called without a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _build

_u64, _i32, _vp, _dbl = C.c_uint64, C.c_int, C.c_void_p, C.c_double

# (name, restype, argtypes) for every symbol include/fock_b200.h declares
SYMBOLS = [
    ("fock_create", _i32, [_i32, C.POINTER(_vp)]),
    ("fock_destroy", _i32, [_vp]),
    ("fock_last_error", C.c_char_p, []),
    ("fock_version", C.c_char_p, []),
    ("fock_device_info", _i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(C.c_size_t)]),
    ("fock_check_status", _i32, [_vp, _vp]),
    ("fock_count", _u64, [_i32, _i32]),
    ("fock_rank_host", _i32, [_i32, _i32, _vp, _u64, _vp]),
    ("fock_unrank_host", _i32, [_i32, _i32, _vp, _u64, _vp]),
    ("fock_rank", _i32, [_vp, _i32, _i32, _vp, _u64, _vp, _vp]),
    ("fock_unrank", _i32, [_vp, _i32, _i32, _vp, _u64, _vp, _vp]),
    ("fock_enumerate", _i32, [_vp, _i32, _i32, _u64, _u64, _vp, _vp]),
    ("fock_mask_match", _i32, [_vp, _i32, _i32, _vp, _i32, _u64, _i32, _u64, _u64, _vp, _vp]),
    ("fock_mask_match_host", _i32, [_i32, _i32, _vp, _i32, _u64, _i32, _vp, _u64, _vp]),
    ("slos_layer", _i32, [_vp, _i32, _i32, _vp, _i32, _vp, _u64, _u64, _vp, _u64, _u64, _vp]),
    ("slos_layer_probs", _i32, [_vp, _i32, _i32, _vp, _i32, _vp, _u64, _u64, _vp, _vp, _vp, _dbl, _u64, _u64, _vp]),
    ("slos_layer_seg", _i32, [_vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _u64, _u64, _vp]),
    ("slos_layer_probs_seg", _i32, [_vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _dbl, _u64, _u64, _vp]),
    ("slos_layer_slab", _i32, [_vp, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp]),
    ("slos_layer_masked", _i32, [_vp, _i32, _i32, _vp, _i32, _vp, _u64, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _dbl, _vp]),
    ("slos_probs_epilogue", _i32, [_vp, _i32, _i32, _vp, _dbl, _vp, _vp, _u64, _u64, _vp]),
    ("slos_amplitudes_epilogue", _i32, [_vp, _i32, _i32, _vp, _dbl, _vp, _u64, _u64, _vp]),
    ("slos_order", _i32, [_i32, _vp, _vp]),
    ("slos_prob_distribution", _i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("slos_prob_distribution_host", _i32, [_vp, _i32, _vp, _vp, _vp, _vp]),
    ("glynn_permanent_batch", _i32, [_vp, _i32, _vp, _u64, _vp, _u64, _u64, _vp]),
    ("glynn_permanent_batch_host", _i32, [_vp, _i32, _vp, _u64, _vp]),
    ("naive_amplitudes", _i32, [_vp, _i32, _i32, _vp, _vp, _vp, _u64, _vp, _vp]),
    ("naive_amplitudes_states", _i32, [_vp, _i32, _i32, _vp, _vp, _vp, _u64, _vp, _vp]),
    ("cc2017_samples", _i32, [_vp, _i32, _i32, _vp, _vp, _u64, _u64, _u64, _vp, _vp]),
    ("cc2017_samples_host", _i32, [_vp, _i32, _i32, _vp, _vp, _u64, _u64, _u64, _vp]),
    ("fock_measure_peak", _i32, [_vp, _i32, C.POINTER(_dbl)]),
    ("fock_profile_events", _i32, [_vp, _vp, _vp]),
    ("fock_launch_count", _u64, [_vp]),
]

_lock = threading.Lock()
_lib = None


class FockError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) libfock_b200.so and declare every prototype."""
    global _lib
    with _lock:
        if _lib is None:
            path = os.environ.get("FOCK_B200_LIB") or _build.LIB     # FOCK_B200_LIB: a tuning build (tools/build_variant.py)
            if path == _build.LIB and (not os.path.exists(path) or os.environ.get("FOCK_B200_REBUILD")):
                _build.build()
            L = C.CDLL(path)
            for name, res, args in SYMBOLS:
                fn = getattr(L, name)  # AttributeError if the library does not export it
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().fock_last_error().decode(errors="replace")
        if rc == -1 and "outside" not in msg and "limit" not in msg:
            raise FockError(f"{what}: {msg} (rc={rc})")
        raise FockError(f"{what}: {msg} (rc={rc})")


_ctx = {}


def context(device: int = 0) -> int:
    """Per-device engine context (opaque handle as int)."""
    L = load()
    with _lock:
        h = _ctx.get(device)
        if h is None:
            out = _vp()
            rc = L.fock_create(device, C.byref(out))
            if rc != 0:
                raise FockError("fock_create: " + L.fock_last_error().decode(errors="replace")
                                + " -- perceval_b200 has no CPU fallback; a B200 (sm_100a) device is required")
            h = out.value
            _ctx[device] = h
    return h


def count(m: int, n: int) -> int:
    return int(load().fock_count(m, n))
