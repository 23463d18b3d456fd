"""Parity of the IDM CUDA kernels (through the C ABI) with the oracle and the frozen live-reference outputs."""
import numpy as np
import pytest
import torch

from conftest import golden, relerr

pytestmark = pytest.mark.gpu


def T64(a, dev):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device=dev)


def I32(a, dev):
    return torch.tensor(np.asarray(a), dtype=torch.int32, device=dev)


def _pack(g):
    K, n = g["p"].shape
    p = g["p"].ravel(); v = g["v"].ravel()
    params = np.concatenate([g["params"][k] for k in range(K)], axis=1)
    off = np.arange(K + 1) * n
    return p, v, params, off, g["head"]


@pytest.mark.parametrize("tier,dtype,tol_s,tol_g", [("fp64", torch.float64, 1e-12, 1e-10), ("fp32", torch.float32, 2e-6, 2e-4)])
def test_step_vs_golden(dev, tier, dtype, tol_s, tol_g):
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("idm_step_" + tier)
    p, v, params, off, head = _pack(g)
    t = lambda a: torch.tensor(a, dtype=dtype, device=dev)
    flags = dhts_b200.Flags(dev)
    tp = t(p).requires_grad_(); tv = t(v).requires_grad_(); th = t(head).requires_grad_()
    np_, nv_, vf = F.idm_step(tp, tv, t(params), I32(off, dev), th, float(g["dt"]), flags, want_flags=True)
    if tier == "fp64":
        assert ((vf.cpu().numpy() & 3) == g["flags"].ravel()).all()      # clip flags identical
    assert relerr(np_.detach().cpu(), g["np"].ravel()) < tol_s and relerr(nv_.detach().cpu(), g["nv"].ravel()) < tol_s
    ((np_ * t(g["g_np"].ravel())).sum() + (nv_ * t(g["g_ns"].ravel())).sum()).backward()
    K, n = g["p"].shape
    # reference returns [n+1] vectors: ghost entry folded back onto the head vehicle / head deltas
    gp = g["g_p"][:, :n].copy(); gs = g["g_s"][:, :n].copy()
    gp[:, -1] += g["g_p"][:, n]; gs[:, -1] += g["g_s"][:, n]
    assert relerr(tp.grad.cpu(), gp.ravel()) < tol_g and relerr(tv.grad.cpu(), gs.ravel()) < tol_g
    gh = np.stack([g["g_p"][:, n], -g["g_s"][:, n]], -1)
    assert relerr(th.grad.cpu(), gh) < tol_g


@pytest.mark.parametrize("ckpt_every", [1, 5, 32])
def test_rollout_vs_golden_fp64(dev, ckpt_every):
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("idm_rollout_fp64")
    T, L, n = int(g["T"]), int(g["L"]), int(g["n"])
    params = np.concatenate([g["params"][l] for l in range(L)], axis=1)
    flags = dhts_b200.Flags(dev)
    tp = T64(g["p0"].ravel(), dev).requires_grad_(); tv = T64(g["v0"].ravel(), dev).requires_grad_()
    th = T64(g["head"], dev).requires_grad_()
    pT, vT = F.idm_rollout(tp, tv, T64(params, dev), I32(np.arange(L + 1) * n, dev), th, float(g["dt"]), T,
                           ckpt_every=ckpt_every, flags=flags)
    ((pT * T64(g["w_p"].ravel(), dev)).sum() + (vT * T64(g["w_v"].ravel(), dev)).sum()).backward()
    flags.check()
    assert relerr(pT.detach().cpu(), g["pT"].ravel()) < 1e-11 and relerr(vT.detach().cpu(), g["vT"].ravel()) < 1e-11
    assert relerr(tp.grad.cpu(), g["g_p0"].ravel()) < 1e-9 and relerr(tv.grad.cpu(), g["g_v0"].ravel()) < 1e-9
    assert relerr(th.grad.cpu(), g["g_head"]) < 1e-9


def test_rollout_vs_golden_fp32(dev):
    """fp32 build vs the reference as shipped, T=300: states 2e-5, gradients 1e-3 of the largest entry."""
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("idm_rollout_fp32")
    T, L, n = int(g["T"]), int(g["L"]), int(g["n"])
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
    params = np.concatenate([g["params"][l] for l in range(L)], axis=1)
    tp = t(g["p0"].ravel()).requires_grad_(); tv = t(g["v0"].ravel()).requires_grad_()
    pT, vT = F.idm_rollout(tp, tv, t(params), I32(np.arange(L + 1) * n, dev), t(g["head"]), float(g["dt"]), T)
    ((pT * t(g["w_p"].ravel())).sum() + (vT * t(g["w_v"].ravel())).sum()).backward()
    assert relerr(pT.detach().cpu(), g["pT"].ravel()) < 2e-5 and relerr(vT.detach().cpu(), g["vT"].ravel()) < 2e-5
    assert relerr(tp.grad.cpu(), g["g_p0"].ravel()) < 1e-3 and relerr(tv.grad.cpu(), g["g_v0"].ravel()) < 1e-3


@pytest.mark.parametrize("sizes,T,K", [([10], 500, 32), ([64] * 9, 40, 16), ([1, 0, 33, 7, 0, 32, 65, 2], 30, 4),
                                        ([130, 5, 256], 12, 3), ([300, 4], 6, 2)])
def test_rollout_ragged_vs_oracle(dev, sizes, T, K):
    """C2's shape (1 x 10 x 500), C5's lane shape (64), ragged lanes incl. empty and single-vehicle lanes,
    multi-slot lanes (>32 vehicles per warp), and a lane too long for the fused kernel (step-kernel path)."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(sum(sizes) + T)
    umax, dt = 30.0, 0.01
    off = np.concatenate([[0], np.cumsum(sizes)]); V, L = int(off[-1]), len(sizes)
    p0 = np.concatenate([np.arange(n) * 15.0 + rng.uniform(0, 5, n) for n in sizes]) if V else np.zeros(0)
    v0 = rng.uniform(3, 21, V)
    par = np.stack([rng.uniform(1.5, 2.0, V) * umax, rng.uniform(1.0, 1.5, V) * umax, rng.uniform(0.8, 1.2, V) * umax,
                    rng.uniform(1, 2, V), rng.uniform(0.2, 0.6, V), np.full(V, 5.0)])
    head = np.stack([rng.uniform(15, 1000, L), rng.uniform(-3, 3, L)], -1)
    wp = rng.normal(size=V); wv = rng.normal(size=V)
    flags = dhts_b200.Flags(dev)
    tp = T64(p0, dev).requires_grad_(); tv = T64(v0, dev).requires_grad_(); th = T64(head, dev).requires_grad_()
    pT, vT = F.idm_rollout(tp, tv, T64(par, dev), I32(off, dev), th, dt, T, ckpt_every=K, flags=flags)
    ((pT * T64(wp, dev)).sum() + (vT * T64(wv, dev)).sum()).backward()
    bits, ncol = flags.check(quiet_collisions=True)
    o = O.idm_rollout(p0, v0, par, off, head, dt, T, g_pT=wp, g_vT=wv)
    assert ncol == o["ncol"]
    assert relerr(pT.detach().cpu(), o["pT"]) < 1e-11 and relerr(vT.detach().cpu(), o["vT"]) < 1e-11
    assert relerr(tp.grad.cpu(), o["g_p0"]) < 1e-9 and relerr(tv.grad.cpu(), o["g_v0"]) < 1e-9
    assert relerr(th.grad.cpu(), o["g_head"]) < 1e-9


def test_collisions_and_clips_vs_oracle(dev):
    """Tight gaps: collisions (print-and-continue convention), acceleration clip, s* clip -- states and flags."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(9)
    n, L, umax, dt = 24, 6, 30.0, 0.01
    V = n * L; off = np.arange(L + 1) * n
    p0 = np.concatenate([np.cumsum(rng.uniform(4.0, 8.0, n)) for _ in range(L)])      # some gaps < vehicle length
    v0 = rng.uniform(0, 25, V); v0[rng.choice(V, 20, replace=False)] = 0.01
    par = np.stack([np.full(V, umax), np.full(V, 0.8 * umax), np.full(V, 0.9 * umax), np.full(V, 0.5), np.full(V, 0.1),
                    np.full(V, 5.0)])
    head = np.tile([[1000.0, 0.0]], (L, 1))
    flags = dhts_b200.Flags(dev)
    np_, nv_, vf = F.idm_step(T64(p0, dev), T64(v0, dev), T64(par, dev), I32(off, dev), T64(head, dev), dt, flags,
                              want_flags=True)
    tot = 0
    for l in range(L):
        s = slice(off[l], off[l + 1])
        o = O.idm_step(p0[s], v0[s], par[:, s], 1000.0, 0.0, dt)
        assert (vf[s].cpu().numpy() == o["flags"]).all()
        assert relerr(nv_[s].cpu(), o["nv"]) < 1e-12
        tot += o["ncol"]
    bits, ncol = flags.read()
    assert tot > 0 and ncol == tot and bits & 4


@pytest.mark.parametrize("n,K", [(64, 16), (20, 1), (130, 8)])
def test_rollout_nan_gradient_flag(dev, n, K):
    """A NaN in the terminal adjoint of one vehicle stays in that vehicle through every adjoint step, so the single test of
    the final adjoint raises the NaN-gradient flag (the reference asserts per step, dmacro_lane.py:308 convention); clean
    inputs leave it clear.  One-, two- and multi-slot warps."""
    import dhts_b200
    from dhts_b200 import functional as F
    from dhts_b200._lib import FLAG_NAN_GRAD
    rng = np.random.default_rng(n)
    umax, dt, T, L = 30.0, 0.01, 40, 3
    V = L * n
    off = np.arange(L + 1) * n
    p0 = np.concatenate([np.arange(n) * 20.0 + rng.uniform(0, 5, n) for _ in range(L)]); v0 = rng.uniform(9, 21, V)
    par = np.stack([np.full(V, umax), np.full(V, 0.8 * umax), np.full(V, 0.9 * umax), np.full(V, 0.5), np.full(V, 0.1), np.full(V, 5.0)])
    head = np.tile(np.array([[1000.0, 0.0]]), (L, 1))
    for poison in (False, True):
        tp = T64(p0, dev).requires_grad_()
        flags = dhts_b200.Flags(dev)
        pT, vT = F.idm_rollout(tp, T64(v0, dev), T64(par, dev), I32(off, dev), T64(head, dev), dt, T, ckpt_every=K, flags=flags)
        w = torch.ones_like(pT)
        if poison:
            w[n + n // 2] = float("nan")
        (pT * w).sum().backward()
        bits, _ = flags.read()
        assert bool(bits & FLAG_NAN_GRAD) == poison
        assert bool(torch.isnan(tp.grad[n:2 * n]).any()) == poison and not bool(torch.isnan(tp.grad[:n]).any())


@pytest.mark.parametrize("L,n,T", [(6, 64, 40), (33, 10, 25), (4, 200, 12)])
def test_rollout_per_step_coupling_vs_oracle(dev, L, n, T):
    """Head deltas given PER STEP (what RoadNetwork.setup_micro_boundary does for a lane inside a network,
    road_network.py:429-580) and a loss that reads the state before every step: final state, gradients wrt the
    initial state and wrt every step's head deltas, against the oracle (pinned to the live reference by
    tests/test_oracle_golden.py; its per-step extension is the composition of terminal-adjoint rollouts)."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(L * 100 + n)
    V, umax, dt = L * n, 30.0, 0.01
    p0 = (np.arange(n)[None] * 16.0 + rng.uniform(0, 6, (L, n))).ravel(); v0 = rng.uniform(4, 22, V)
    par = np.stack([rng.uniform(1.5, 2, V) * umax, rng.uniform(1, 1.5, V) * umax, rng.uniform(0.8, 1.2, V) * umax,
                    rng.uniform(1, 2, V), rng.uniform(0.2, 0.6, V), np.full(V, 5.0)])
    off = np.arange(L + 1) * n
    head = np.stack([rng.uniform(15, 80, (T, L)), rng.uniform(-3, 3, (T, L))], -1)
    g_hist = rng.normal(size=(T, V, 2)); wp = rng.normal(size=V); wv = rng.normal(size=V)
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    flags = dhts_b200.Flags(dev)
    tp, tv, th_ = t(p0).requires_grad_(), t(v0).requires_grad_(), t(head).requires_grad_()
    pT, vT, ph, vh = F.idm_rollout(tp, tv, t(par), torch.tensor(off, dtype=torch.int32, device=dev), th_, dt, T, ckpt_every=1,
                                   flags=flags, return_history=True)
    ((ph * t(g_hist[..., 0])).sum() + (vh * t(g_hist[..., 1])).sum() + (pT * t(wp)).sum() + (vT * t(wv)).sum()).backward()
    flags.check(quiet_collisions=True)
    o = O.idm_rollout(p0, v0, par, off, head, dt, T, g_pT=wp, g_vT=wv, g_hist=g_hist, want_hist=True)
    assert relerr(pT.detach().cpu(), o["pT"]) < 1e-10 and relerr(vT.detach().cpu(), o["vT"]) < 1e-10
    assert relerr(ph.detach().cpu(), o["hist"][:T, :, 0]) < 1e-10
    assert relerr(tp.grad.cpu(), o["g_p0"]) < 1e-9 and relerr(tv.grad.cpu(), o["g_v0"]) < 1e-9
    assert relerr(th_.grad.cpu(), o["g_head"]) < 1e-9
