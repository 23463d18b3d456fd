"""Importable alias of the product package.

The product lives in ``diff-hybrid-traffic-sim_b200/`` (the directory name the
project layout prescribes); a hyphenated name cannot be imported, so this tiny
package points its ``__path__`` at that directory and runs its ``__init__``.
``import dhts_b200`` / ``from dhts_b200 import ops`` is the supported spelling.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "diff-hybrid-traffic-sim_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
