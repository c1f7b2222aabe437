"""sw_solver_b200 -- B200-native drop-in for the time-stepping path of ai2cm/sw_solver.

Mirrors the reference package surface (``src/sw_solver/__init__.py:3-5``): the solver
module (here ``sw_solver_b200.solver``, also reachable as ``sw_solver_b200.b200`` and, for
code written against the reference, ``sw_solver_b200.numpy``), ``CartesianGrid``,
``LatLonGrid`` and ``ICType``.
"""

from . import grid, ic, utils  # noqa: F401
from .grid import CartesianGrid, LatLonGrid  # noqa: F401
from .ic import ICType  # noqa: F401


def __getattr__(name):
    # the solver needs torch + the CUDA library; import it on first use so that grid / IC
    # utilities stay importable on their own
    if name in ("solver", "b200", "numpy"):
        import importlib

        mod = importlib.import_module(".solver", __name__)
        globals()["solver"] = globals()["b200"] = globals()["numpy"] = mod
        return mod
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
