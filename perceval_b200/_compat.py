"""Binds the host side either to a real Perceval install or to the local mirrors.

Conformance with Perceval requires real subclassing (``Processor`` asserts ``isinstance(backend, ABackend)``,
reference perceval/components/processor.py:119,165-166), so when ``import perceval`` works the device backends
derive from Perceval's own ABCs and return exqalibur state types.  In this image exqalibur cannot be installed
(SURVEY.md 0.2), so the structurally identical mirrors in ``perceval_b200.backends._abstract`` / ``.states`` are used.
"""
from __future__ import annotations

import os

HAVE_PERCEVAL = False
if not os.environ.get("PERCEVAL_B200_STANDALONE"):
    try:  # pragma: no cover - not importable in the build image
        import perceval as _pcvl  # noqa: F401
        from perceval.backends import ABackend, ASamplingBackend, AStrongSimulationBackend  # noqa: F401
        from perceval.utils import BasicState, BSDistribution, BSSamples, StateVector  # noqa: F401
        from exqalibur import FockState  # noqa: F401
        try:
            from perceval.utils import BSCount  # noqa: F401
        except Exception:
            from .states import BSCount  # noqa: F401
        HAVE_PERCEVAL = True
    except Exception:
        HAVE_PERCEVAL = False

if not HAVE_PERCEVAL:
    from .states import BasicState, BSCount, BSDistribution, BSSamples, FockState, StateVector  # noqa: F401
    from .backends._abstract import ABackend, ASamplingBackend, AStrongSimulationBackend  # noqa: F401

MIN_P = 1e-16  # perceval/utils/globals.py:30-34 global_params["min_p"]
if HAVE_PERCEVAL:  # pragma: no cover
    try:
        from perceval.utils import global_params as _gp
        MIN_P = _gp["min_p"]
    except Exception:
        pass
