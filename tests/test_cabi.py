"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/dhts.h declares
(no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "dhts.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dhts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import dhts_b200
    so = dhts_b200.build()
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dhts_version() >= 100
    # the Python binding's symbol table and the header agree
    assert sorted(dhts_b200._lib.SYMBOLS) == names


def test_no_cpu_fallback_without_cuda_tensors():
    import pytest
    import torch
    import dhts_b200
    from dhts_b200 import functional as F
    with pytest.raises(RuntimeError, match="CUDA only"):
        dhts_b200._lib.require_cuda(torch.zeros(3))
    r = torch.rand(2, 8, dtype=torch.float64)
    with pytest.raises(Exception):
        F.arz_rollout(r, r * 30, r[:, :2], r[:, :2] * 30, 5.0, 30.0, 0.01, 2)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "diff-hybrid-traffic-sim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "_lib.py" and "oracle" not in txt, (dirpath, f)
                assert "/root/reference" not in txt, (dirpath, f)
