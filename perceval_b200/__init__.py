"""perceval_b200 -- B200-native Fock-amplitude engine behind Perceval's backend API.

Provides SLOS_B200, Naive_B200 and CliffordClifford2017_B200 (drop-in ABackend subclasses) on top of the C ABI in
include/fock_b200.h (libfock_b200.so, hand-written CUDA for sm_100a, complex128).  Importing the package does not
touch the GPU; the first compute call does, and raises if there is none (no CPU fallback).
"""
from ._lib import FockError, lib_path, load  # noqa: F401
from .backends import (BACKEND_LIST, B200_BACKENDS, BackendFactory, Clifford2017B200Backend, NaiveB200Backend,  # noqa: F401
                       SLOSB200Backend, register, set_seed)
from .circuit import UnitaryCircuit, random_unitary  # noqa: F401
from ._compat import HAVE_PERCEVAL, BasicState, BSDistribution, BSSamples, FockState, StateVector  # noqa: F401

__version__ = "0.1.0"
