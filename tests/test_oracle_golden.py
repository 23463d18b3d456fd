"""Pins the CPU oracle (oracle/dhts_oracle.c) against outputs of the LIVE reference frozen in
tests/golden/ by oracle/gen_golden.py: dtype-proxied fp64 reference (tier 2) to rounding, and the
reference as shipped (fp32 storage, tier 1) to fp32 rounding."""
import numpy as np
import pytest

from conftest import golden, relerr
from oracle import oracle as O

TIERS = [("fp64", False, 1e-12, 1e-11), ("fp32", True, 2e-5, 2e-4)]


@pytest.mark.parametrize("tier,f32,tol_s,tol_g", TIERS)
def test_arz_step(tier, f32, tol_s, tol_g):
    g = golden("arz_step_" + tier)
    hist = np.zeros(3, dtype=int)
    for k in range(g["pr"].shape[0]):
        o = O.arz_step(g["pr"][k], g["py"][k], g["pu"][k], g["pe"][k], float(g["dx"]), float(g["umax"]), float(g["dt"]),
                       f32=f32)
        assert (o["case"] == g["case"][k]).all()
        hist += np.bincount(o["case"], minlength=3)
        assert np.abs(o["speeds"] - g["speeds"][k]).max() < (1e-9 if not f32 else 1e-3)
        assert relerr(o["nr"], g["nr"][k]) < tol_s and relerr(o["ny"], g["ny"][k]) < tol_s
        assert relerr(o["nu"], g["nu"][k]) < max(tol_s, 1e-12) * 10
        assert relerr(o["dqs"], g["dqs"][k]) < tol_g
        gr, gy = O.arz_vjp(o["dqs"], g["g_nr"][k], g["g_ny"][k], f32=f32)
        assert relerr(gr, g["g_r"][k]) < tol_g and relerr(gy, g["g_y"][k]) < tol_g
        assert o["cfl"] == 0
    assert (hist > 50).all(), hist      # all three Riemann outcomes are exercised


@pytest.mark.parametrize("tier,f32,tol_s,tol_g", TIERS)
def test_arz_rollout(tier, f32, tol_s, tol_g):
    g = golden("arz_rollout_" + tier)
    T = int(g["T"])
    o = O.arz_rollout(g["r0"], g["u0"], g["ghost_ru"], float(g["dx"]), float(g["umax"]), float(g["dt"]), T, f32=f32,
                      g_rT=g["w_r"], g_uT=g["w_u"], want_hist=True)
    assert o["cfl"] == 0
    for a, b in (("rT", "rT"), ("yT", "yT"), ("uT", "uT")):
        assert relerr(o[a], g[b]) < tol_s * 10, a
    for i, t in enumerate(g["snaps"]):
        assert relerr(o["hist"][int(t)], g["snap"][:, i]) < tol_s * 10
    for a in ("g_r0", "g_u0", "g_ghost"):
        assert relerr(o[a], g[a]) < tol_g, a


def test_arz_rollout_with_vacuum():
    """Lanes with empty / near-vacuum stretches and vacuum ghosts, 60 steps of the live reference (oracle/gen_golden_vac.py):
    the vacuum branches of the case tree and the below-eps fix-ups of u_eq / flux_prime in the oracle's rollout and adjoint."""
    g = golden("arz_rollout_vac_fp64")
    o = O.arz_rollout(g["r0"], g["u0"], g["ghost_ru"], float(g["dx"]), float(g["umax"]), float(g["dt"]), int(g["T"]),
                      g_rT=g["w_r"], g_uT=g["w_u"])
    assert o["cfl"] == 0 and (g["rT"] < 1e-5).any() and (g["r0"] == 0).any()
    for a in ("rT", "yT", "uT"):
        assert relerr(o[a], g[a]) < 1e-12, a
    for a in ("g_r0", "g_u0", "g_ghost"):
        assert relerr(o[a], g[a]) < 1e-11, a


@pytest.mark.parametrize("tier,f32,tol_s,tol_g", TIERS)
def test_idm_step(tier, f32, tol_s, tol_g):
    g = golden("idm_step_" + tier)
    seen = 0
    for k in range(g["p"].shape[0]):
        o = O.idm_step(g["p"][k], g["v"][k], g["params"][k], g["head"][k][0], g["head"][k][1], float(g["dt"]), f32=f32)
        assert ((o["flags"] & 3) == g["flags"][k]).all()
        seen |= int(np.bitwise_or.reduce(o["flags"]))
        assert relerr(o["np"], g["np"][k]) < tol_s and relerr(o["nv"], g["nv"][k]) < tol_s
        assert relerr(o["dqs"], g["dqs"][k]) < tol_g
        gp, gs = O.idm_vjp(o["dqs"], g["g_np"][k], g["g_ns"][k], f32=f32)
        assert relerr(gp, g["g_p"][k]) < tol_g and relerr(gs, g["g_s"][k]) < tol_g
    assert seen & 1 and seen & 2        # both clips are exercised


@pytest.mark.parametrize("tier,f32,tol_s,tol_g", TIERS)
def test_idm_rollout(tier, f32, tol_s, tol_g):
    g = golden("idm_rollout_" + tier)
    T, L, n = int(g["T"]), int(g["L"]), int(g["n"])
    params = np.concatenate([g["params"][l] for l in range(L)], axis=1)
    off = np.arange(L + 1) * n
    o = O.idm_rollout(g["p0"].ravel(), g["v0"].ravel(), params, off, g["head"], float(g["dt"]), T, f32=f32,
                      g_pT=g["w_p"].ravel(), g_vT=g["w_v"].ravel(), want_hist=True)
    assert o["ncol"] == 0
    assert relerr(o["pT"], g["pT"].ravel()) < tol_s and relerr(o["vT"], g["vT"].ravel()) < tol_s
    for i, t in enumerate(g["snaps"]):
        assert relerr(o["hist"][int(t)].reshape(L, n, 2), g["snap"][:, i]) < tol_s
    assert relerr(o["g_p0"], g["g_p0"].ravel()) < tol_g and relerr(o["g_v0"], g["g_v0"].ravel()) < tol_g
    assert relerr(o["g_head"], g["g_head"]) < tol_g


def test_oracle_edge_cases():
    # one-cell lane, vacuum everywhere, collision flagging, empty lanes
    o = O.arz_step([0.3, 0.5, 0.2], [0.0, 0.1, 0.0], [10.0, 12.0, 9.0], [20.0, 18.0, 22.0], 5.0, 30.0, 0.01)
    assert o["nr"].shape == (1,) and np.isfinite(o["nr"]).all()
    o = O.arz_step(np.zeros(6), np.zeros(6), np.full(6, 30.0), np.full(6, 30.0), 5.0, 30.0, 0.01)
    assert (o["case"] == 0).all() and np.abs(o["nr"]).max() == 0
    assert O.arz_step([0.5, 0.5, 0.5], [0, 0, 0], [29.0, 29.0, 29.0], [9.0, 9.0, 9.0], 0.1, 30.0, 0.01)["cfl"] == 1
    par = np.array([[30.0] * 2, [24.0] * 2, [27.0] * 2, [0.5] * 2, [0.1] * 2, [5.0] * 2])
    o = O.idm_step([0.0, 3.0], [10.0, 10.0], par, 1000.0, 0.0, 0.01)
    assert o["ncol"] == 1 and o["flags"][0] & 4
    o = O.idm_rollout(np.zeros(0), np.zeros(0), np.zeros((6, 0)), [0, 0, 0], np.zeros((2, 2)), 0.01, 3,
                      g_pT=np.zeros(0), g_vT=np.zeros(0))
    assert o["pT"].shape == (0,)


def test_arz_per_step_coupling_vs_live_reference():
    """Per-step ghosts and a loss that reads the state before every step (oracle/gen_golden_perstep.py, live fp64 reference):
    the oracle's g_hist injection and per-step ghost adjoints."""
    from oracle import oracle as O
    g = golden("arz_perstep_fp64")
    T, B, N, umax = int(g["T"]), int(g["B"]), int(g["N"]), float(g["umax"])
    f = O.arz_rollout(g["r0"], g["u0"], g["ghost_ru"], float(g["dx"]), umax, float(g["dt"]), T, want_hist=True)
    hist = f["hist"]                                       # [T+1,B,N,2]
    assert relerr(hist[:T, :, :, 0], g["r_hist"]) < 1e-12 and relerr(f["rT"], g["rT"]) < 1e-12 and relerr(f["uT"], g["uT"]) < 1e-12
    # the loss reads (r, u) of the state before step t: chain u = compute_u(r, y) into (r, y) adjoints
    g_hist = np.zeros((T, B, N, 2))
    for t in range(T):
        for b in range(B):
            for j in range(N):
                dr, dy = _du_dry(hist[t, b, j, 0], hist[t, b, j, 1], umax)
                g_hist[t, b, j, 0] = g["w_r"][t, b, j] + g["w_u"][t, b, j] * dr
                g_hist[t, b, j, 1] = g["w_u"][t, b, j] * dy
    o = O.arz_rollout(g["r0"], g["u0"], g["ghost_ru"], float(g["dx"]), umax, float(g["dt"]), T, g_rT=g["wT_r"], g_uT=g["wT_u"],
                      g_hist=g_hist)
    assert relerr(o["g_r0"], g["g_r0"]) < 1e-10 and relerr(o["g_u0"], g["g_u0"]) < 1e-10
    assert relerr(o["g_ghost"], g["g_ghost"]) < 1e-10


def _du_dry(r, y, umax, eps=1e-5):
    """d compute_u / d(r, y), the true derivative autograd takes outside the operator (model/macro/_arz.py:126-138)."""
    if r >= eps:
        return -y / (r * r) - 0.5 * umax / np.sqrt(r + eps), 1.0 / r
    return 0.0, 1.0 / eps
