"""Device backends + the registry / factory surface of reference perceval/backends/__init__.py:39-62."""
from __future__ import annotations

import warnings

from .._compat import HAVE_PERCEVAL, ABackend, ASamplingBackend, AStrongSimulationBackend
from .clifford2017 import Clifford2017B200Backend, set_seed
from .naive import NaiveB200Backend
from .slos import SLOSB200Backend

B200_BACKENDS = {
    "SLOS_B200": SLOSB200Backend,
    "Naive_B200": NaiveB200Backend,
    "CliffordClifford2017_B200": Clifford2017B200Backend,
}


def register() -> dict:
    """Plug the device backends into Perceval's registry.  ``Processor("SLOS_B200", ...)``,
    ``BackendFactory.get_backend("SLOS_B200")`` and ``SimulatorFactory.build(..., backend="SLOS_B200")`` all read the
    same mutable dict (reference perceval/backends/__init__.py:39, components/processor.py:112-116,
    simulators/simulator_factory.py:111-115), so this assignment is the whole plug-in mechanism."""
    if HAVE_PERCEVAL:  # pragma: no cover - Perceval is not importable in the build image
        from perceval.backends import BACKEND_LIST as PCVL_LIST
        PCVL_LIST.update(B200_BACKENDS)
        return PCVL_LIST
    return BACKEND_LIST


if HAVE_PERCEVAL:  # pragma: no cover
    from perceval.backends import BACKEND_LIST, BackendFactory
    register()
else:
    # stand-alone mirror: the reference names resolve to the device implementations so that code written against
    # BackendFactory.get_backend("SLOS") keeps working unchanged
    BACKEND_LIST = dict(B200_BACKENDS)
    BACKEND_LIST.update({"SLOS": SLOSB200Backend, "Naive": NaiveB200Backend, "CliffordClifford2017": Clifford2017B200Backend})

    class BackendFactory:
        @staticmethod
        def get_backend(backend_name: str = "SLOS", **kwargs) -> ABackend:
            if backend_name in BACKEND_LIST:
                return BACKEND_LIST[backend_name](**kwargs)
            warnings.warn(f'Backend "{backend_name}" not found. Falling back on SLOS')
            return BACKEND_LIST["SLOS"](**kwargs)

        @staticmethod
        def list():
            return list(BACKEND_LIST.keys())

__all__ = ["ABackend", "ASamplingBackend", "AStrongSimulationBackend", "SLOSB200Backend", "NaiveB200Backend",
           "Clifford2017B200Backend", "BACKEND_LIST", "BackendFactory", "register", "set_seed", "B200_BACKENDS"]
