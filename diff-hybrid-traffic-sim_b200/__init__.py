"""B200-native simulation step of diff-hybrid-traffic-sim: ARZ macro lanes, IDM micro
lanes, the macro<->micro exchange and their adjoints, as hand-written sm_100a CUDA
kernels behind a C ABI (include/dhts.h), driven from torch.autograd.Functions.

Import as ``dhts_b200`` (the directory name ``diff-hybrid-traffic-sim_b200`` is not a
valid Python identifier; ``dhts_b200/`` is a path alias onto it).
"""
from . import _lib                      # noqa: F401
from ._build import build               # noqa: F401
from ._lib import Flags, UnsupportedShape  # noqa: F401


def __getattr__(name):
    # torch-facing modules are imported lazily so that `build()` works before the .so exists
    if name in ("ops", "functional", "dist", "dropin", "network", "itscp", "hybrid_network", "itscp_env", "control",
                "inverse", "run_itscp", "run_inverse"):
        import importlib
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
