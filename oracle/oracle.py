"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front-end of the CPU oracle.

The oracle is the checker for the CUDA product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under
``diff-hybrid-traffic-sim_b200/`` does (tests/test_layout.py greps for it).

The arithmetic lives in ``oracle/dhts_oracle.c`` (each function cites the
reference file:line it restates).  This file only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdhts_oracle.so")
_SRC = os.path.join(_HERE, "dhts_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle.  $CC in this image points at a gcc without
    libgomp, so probe the system compilers and fall back to no OpenMP."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    base = ["-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-shared", "-o", _SO, _SRC, "-lm"]
    tried = []
    for cc in ("/usr/bin/gcc", shutil.which("gcc"), shutil.which("cc")):
        if not cc or cc in tried:
            continue
        tried.append(cc)
        for omp in (["-fopenmp"], []):
            r = subprocess.run([cc] + omp + base, capture_output=True, text=True)
            if r.returncode == 0:
                return _SO
    raise RuntimeError("could not build the oracle with any of %s" % tried)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_u_eq.restype = ctypes.c_double
        _lib.orc_u_eq.argtypes = [ctypes.c_double] * 2
        _lib.orc_compute_u.restype = ctypes.c_double
        _lib.orc_compute_u.argtypes = [ctypes.c_double] * 3
        _lib.orc_compute_y.restype = ctypes.c_double
        _lib.orc_compute_y.argtypes = [ctypes.c_double] * 3
    return _lib


def set_threads(n: int) -> int:
    """OpenMP threads of the lane-parallel rollouts; returns the count in effect (torchrun exports OMP_NUM_THREADS=1)."""
    return int(lib().orc_set_threads(int(n)))


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def u_eq(r, umax):
    return np.vectorize(lambda x: lib().orc_u_eq(float(x), float(umax)))(np.asarray(r, dtype=np.float64))


def compute_u(r, y, umax):
    f = lib().orc_compute_u
    return np.vectorize(lambda a, b: f(float(a), float(b), float(umax)))(_d(r), _d(y))


def compute_y(r, u, umax):
    f = lib().orc_compute_y
    return np.vectorize(lambda a, b: f(float(a), float(b), float(umax)))(_d(r), _d(u))


def arz_step(pad_r, pad_y, pad_u, pad_ueq, dx, umax, dt, f32=False, want_jac=True):
    """One lane, one step.  pad_* are [N+2] (ghost, cells, ghost).
    Returns dict(nr, ny, nu, case, speeds, dqs, cfl)."""
    pad_r, pad_y, pad_u, pad_ueq = map(_d, (pad_r, pad_y, pad_u, pad_ueq))
    N = pad_r.shape[0] - 2
    nr = np.zeros(N); ny = np.zeros(N); nu = np.zeros(N)
    case = np.zeros(N + 1, dtype=np.int32)
    speeds = np.zeros((N + 1, 2))
    dqs = np.zeros((N, 3, 2, 2)) if want_jac else None
    cfl = lib().orc_arz_step(_p(pad_r), _p(pad_y), _p(pad_u), _p(pad_ueq), ctypes.c_int(N),
                             ctypes.c_double(dx), ctypes.c_double(umax), ctypes.c_double(dt),
                             ctypes.c_int(int(f32)), _p(nr), _p(ny), _p(nu), _p(case), _p(speeds), _p(dqs))
    return dict(nr=nr, ny=ny, nu=nu, case=case, speeds=speeds, dqs=dqs, cfl=int(cfl))


def arz_vjp(dqs, g_nr, g_ny, f32=False):
    dqs = _d(dqs); g_nr = _d(g_nr); g_ny = _d(g_ny)
    N = g_nr.shape[0]
    g_r = np.zeros(N + 2); g_y = np.zeros(N + 2)
    lib().orc_arz_vjp(_p(dqs), ctypes.c_int(N), _p(g_nr), _p(g_ny), ctypes.c_int(int(f32)), _p(g_r), _p(g_y))
    return g_r, g_y


def arz_rollout(r0, u0, ghost_ru, dx, umax, dt, T, f32=False, g_rT=None, g_yT=None, g_uT=None, want_hist=False, g_hist=None):
    """B lanes x N cells.  r0,u0 [B,N]; ghost_ru [B,2,2] static ghosts or [T,B,2,2] one pair per step.
    dx, umax scalars or [B].  g_hist [T,B,N,2]: dLoss/d(r,y) of the state before every step (per-step loss).
    Returns dict(rT,yT,uT,cfl[,g_r0,g_u0,g_ghost][,hist]); g_ghost is [B,2,2] or, with per-step ghosts, [T,B,2,2]."""
    r0 = _d(r0); u0 = _d(u0); ghost_ru = _d(ghost_ru)
    B, N = r0.shape
    tv = ghost_ru.ndim == 4
    if tv:
        assert ghost_ru.shape == (T, B, 2, 2)
    dx = _d(np.broadcast_to(np.asarray(dx, dtype=np.float64), (B,)))
    umax = _d(np.broadcast_to(np.asarray(umax, dtype=np.float64), (B,)))
    rT = np.zeros((B, N)); yT = np.zeros((B, N)); uT = np.zeros((B, N))
    hist = np.zeros((T + 1, B, N, 2)) if want_hist else None
    bwd = g_rT is not None or g_uT is not None or g_yT is not None or g_hist is not None
    out = {}
    if bwd:
        g_rT = _d(g_rT) if g_rT is not None else np.zeros((B, N))
        g_yT = _d(g_yT) if g_yT is not None else np.zeros((B, N))
        g_uT = _d(g_uT) if g_uT is not None else np.zeros((B, N))
        g_hist = _d(g_hist) if g_hist is not None else None
        g_r0 = np.zeros((B, N)); g_u0 = np.zeros((B, N)); g_gh = np.zeros((T, B, 2, 2) if tv else (B, 2, 2))
    else:
        g_r0 = g_u0 = g_gh = None
    cfl = lib().orc_arz_rollout_ex(_p(r0), _p(u0),
                                   _p(ghost_ru) if not tv else None, _p(ghost_ru) if tv else None,
                                   ctypes.c_int(B), ctypes.c_int(N), _p(dx), _p(umax),
                                   ctypes.c_double(dt), ctypes.c_int(T), ctypes.c_int(int(f32)), _p(rT), _p(yT), _p(uT),
                                   _p(g_rT) if bwd else None, _p(g_yT) if bwd else None, _p(g_uT) if bwd else None,
                                   _p(g_hist) if bwd else None, _p(g_r0), _p(g_u0), _p(g_gh), _p(hist))
    out.update(rT=rT, yT=yT, uT=uT, cfl=int(cfl))
    if bwd:
        out.update(g_r0=g_r0, g_u0=g_u0, g_ghost=g_gh)
    if want_hist:
        out["hist"] = hist
    return out


def idm_step(p, v, params, head_dp, head_dv, dt, f32=False, want_jac=True):
    """One lane, one step.  p,v [n]; params [6,n] (a_max,a_pref,v_target,min_space,time_pref,length)."""
    p = _d(p); v = _d(v); params = _d(params)
    n = p.shape[0]
    np_ = np.zeros(n); nv_ = np.zeros(n)
    flags = np.zeros(n, dtype=np.int32)
    dqs = np.zeros((n, 2, 2, 2)) if want_jac else None
    ncol = lib().orc_idm_step(_p(p), _p(v), _p(params), ctypes.c_int(n), ctypes.c_double(head_dp),
                              ctypes.c_double(head_dv), ctypes.c_double(dt), ctypes.c_int(int(f32)), _p(np_), _p(nv_),
                              _p(flags), _p(dqs))
    return dict(np=np_, nv=nv_, flags=flags, dqs=dqs, ncol=int(ncol))


def idm_vjp(dqs, g_np, g_ns, f32=False):
    dqs = _d(dqs); g_np = _d(g_np); g_ns = _d(g_ns)
    n = g_np.shape[0]
    g_p = np.zeros(n + 1); g_s = np.zeros(n + 1)
    lib().orc_idm_vjp(_p(dqs), ctypes.c_int(n), _p(g_np), _p(g_ns), ctypes.c_int(int(f32)), _p(g_p), _p(g_s))
    return g_p, g_s


def idm_rollout(p0, v0, params, lane_off, head, dt, T, f32=False, g_pT=None, g_vT=None, want_hist=False, g_hist=None):
    """CSR lanes.  p0,v0 [V]; params [6,V]; lane_off [L+1] int32; head [L,2] or, one pair per step, [T,L,2].
    g_hist [T,V,2]: dLoss/d(p,v) of the state before every step.  g_head comes back [L,2] or [T,L,2]."""
    p0 = _d(p0); v0 = _d(v0); params = _d(params); head = _d(head)
    lane_off = np.ascontiguousarray(lane_off, dtype=np.int32)
    V = p0.shape[0]; L = lane_off.shape[0] - 1
    tv = head.ndim == 3
    pT = np.zeros(V); vT = np.zeros(V)
    hist = np.zeros((T + 1, V, 2)) if want_hist else None
    bwd = g_pT is not None or g_hist is not None
    if bwd:
        g_pT = _d(g_pT) if g_pT is not None else np.zeros(V); g_vT = _d(g_vT) if g_vT is not None else np.zeros(V)
        g_hist = _d(g_hist) if g_hist is not None else None
        g_p0 = np.zeros(V); g_v0 = np.zeros(V); g_head = np.zeros((T, L, 2) if tv else (L, 2))
    else:
        g_p0 = g_v0 = g_head = None
    ncol = lib().orc_idm_rollout_ex(_p(p0), _p(v0), _p(params), ctypes.c_int(V), _p(lane_off), ctypes.c_int(L),
                                    None if tv else _p(head), _p(head) if tv else None,
                                    ctypes.c_double(dt), ctypes.c_int(T), ctypes.c_int(int(f32)), _p(pT), _p(vT),
                                    _p(g_pT) if bwd else None, _p(g_vT) if bwd else None, _p(g_hist) if bwd else None,
                                    _p(g_p0), _p(g_v0), _p(g_head), _p(hist))
    out = dict(pT=pT, vT=vT, ncol=int(ncol))
    if bwd:
        out.update(g_p0=g_p0, g_v0=g_v0, g_head=g_head)
    if want_hist:
        out["hist"] = hist
    return out
