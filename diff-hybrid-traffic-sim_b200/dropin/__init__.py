"""Drop-in replacement of the reference's ``road`` / ``model`` / ``dmath`` packages.

The reference's callers (``example/inverse/*.py``, ``example/control/itscp/*``) import
``road.lane.dmacro_lane.dMacroLane``, ``road.network.road_network.RoadNetwork`` ... and drive
them through plain attribute access (SURVEY.md 8b lists the surface).  This directory holds
packages of the same import paths whose lanes keep the object API but whose step -- the
body of ``dMacroForwardLayer`` / ``dMicroForwardLayer`` and of ``Conversion.*`` -- runs in
the sm_100a kernels of libdhts_b200.so.  There is no CPU path: stepping a lane without a
CUDA device raises.

    import dhts_b200.dropin as dropin
    dropin.install()                       # puts these packages first on sys.path
    from road.lane.dmacro_lane import dMacroLane      # ours, not the reference's

``install(precision=...)`` selects the arithmetic: ``"mixed"`` (default) stores fp32 state
and evaluates every step in fp64 exactly like the reference as shipped,
``"float64"`` / ``"float32"`` run everything in one precision.
"""
from __future__ import annotations

import os
import sys

from . import runtime  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_TOP = ("road", "model", "dmath")


def install(precision: str | None = None, device=None) -> str:
    """Make ``import road`` / ``model`` / ``dmath`` resolve to this directory.  Modules of the
    same names that were imported from elsewhere (the reference checkout) are evicted."""
    root = os.path.dirname(os.path.dirname(_HERE))
    if root not in sys.path:
        sys.path.insert(0, root)            # so that the drop-in modules can `import dhts_b200`
    if _HERE in sys.path:
        sys.path.remove(_HERE)
    sys.path.insert(0, _HERE)
    for name in list(sys.modules):
        if name.split(".")[0] in _TOP:
            f = getattr(sys.modules[name], "__file__", None) or ""
            if not os.path.abspath(f).startswith(_HERE):
                del sys.modules[name]
    if precision is not None or device is not None:
        runtime.configure(precision=precision, device=device)
    return _HERE


def uninstall() -> None:
    if _HERE in sys.path:
        sys.path.remove(_HERE)
    for name in list(sys.modules):
        if name.split(".")[0] in _TOP:
            del sys.modules[name]
