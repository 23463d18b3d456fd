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


def test_ckpt_size_query_is_host_side_and_matches_the_documented_layout():
    """dhts_arz_rollout_ckpt_elems_* (include/dhts.h) plans on the host: the S stored states [S][2][B][N] and, in ckpt_mode 1
    (every state stored, whole warps per lane), N / 4 bytes of interface-outcome ballots per lane and step behind them."""
    import dhts_b200
    lib = ctypes.CDLL(dhts_b200.build())
    f = lib.dhts_arz_rollout_ckpt_elems_f64
    f.restype = ctypes.c_longlong
    mode = ctypes.c_int(-1)
    B, N, T = 6560, 1024, 1000
    n = f(B, N, T, 1, ctypes.byref(mode))
    assert mode.value == 1 and n == T * 2 * B * N + T * B * (N // 4) // 8        # the bench's lane chunk
    n = f(B, N, T, 32, ctypes.byref(mode))
    assert mode.value == 0 and n == ((T + 31) // 32) * 2 * B * N                  # sparse checkpoints: states only
    n = f(37, 10, 60, 1, ctypes.byref(mode))
    assert mode.value == 0 and n == 60 * 2 * 37 * 10                              # one cell per thread: no outcome storage
    g = lib.dhts_arz_rollout_ckpt_elems_f32
    g.restype = ctypes.c_longlong
    n = g(8, 1024, 10, 1, ctypes.byref(mode))
    assert mode.value == 1 and n == 10 * 2 * 8 * 1024 + 10 * 8 * (1024 // 4) // 4
    assert f(-1, 8, 1, 1, ctypes.byref(mode)) < 0
