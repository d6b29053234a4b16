"""Generates the committed golden fixtures from the CPU oracle (the reference itself cannot be imported here:
exqalibur is not installable offline).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))

# BASELINE config 1: SLOS 6 photons / 12 modes, Haar-random unitary (Matrix.random_unitary restated), seed 0
u = oracle.random_unitary(12, seed=0)
in_state = np.array([1] * 6 + [0] * 6, dtype=np.uint8)
np.savez_compressed(os.path.join(here, "slos_6_12_seed0.npz"), u=u, in_state=in_state,
                    probs=oracle.slos_probs(u, in_state, scatter=True), coefs=oracle.slos_coefs(u, in_state, scatter=True))

# permanents of Haar sub-matrices (config 2 shape) at oracle-friendly sizes
mats, perms = [], []
for n, seed in [(8, 0), (12, 1), (16, 2), (20, 3)]:
    uu = oracle.random_unitary(2 * n, seed=seed)
    mat = np.ascontiguousarray(uu[:n, :n])
    np.savez_compressed(os.path.join(here, f"perm_{n}_seed{seed}.npz"), mat=mat, perm=np.array(oracle.permanent(mat)),
                        ryser=np.array(oracle.permanent_ryser(mat)) if n <= 16 else np.array(np.nan))

# Clifford & Clifford: first 256 samples for n=5, m=10, seed 42
u = oracle.random_unitary(10, seed=2)
st = np.array([1] * 5 + [0] * 5, dtype=np.uint8)
np.savez_compressed(os.path.join(here, "cc2017_5_10_seed42.npz"), u=u, in_state=st,
                    samples=oracle.cc2017_samples(u, st, 256, seed=42, offset=0))
print("golden fixtures written to", here)
