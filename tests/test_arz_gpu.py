"""Parity of the ARZ CUDA kernels (through the C ABI) with the oracle and with the frozen outputs of
the live reference.  Tolerances: fp64 build rtol 1e-5 per the north star (observed ~1e-13; asserted
at 1e-9 so that a real regression is caught); fp32 build stated per test."""
import numpy as np
import pytest
import torch

from conftest import golden, relerr

pytestmark = pytest.mark.gpu


def T64(a, dev):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device=dev)


def test_step_vs_golden_fp64(dev):
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("arz_step_fp64")
    K = g["pr"].shape[0]
    flags = dhts_b200.Flags(dev)
    r_pad = T64(g["pr"], dev).requires_grad_(); y_pad = T64(g["py"], dev).requires_grad_()
    nr, ny, nu, case = F.arz_step(r_pad, y_pad, T64(g["pu"], dev), float(g["dx"]), float(g["umax"]), float(g["dt"]),
                                  flags, ueq_pad=T64(g["pe"], dev), want_case=True)
    assert (case.cpu().numpy() == g["case"]).all()          # every Riemann outcome identical
    assert relerr(nr.detach().cpu(), g["nr"]) < 1e-12 and relerr(ny.detach().cpu(), g["ny"]) < 1e-12
    assert relerr(nu.detach().cpu(), g["nu"]) < 1e-11
    (nr * T64(g["g_nr"], dev)).sum().backward(retain_graph=True)
    gr1 = r_pad.grad.clone(); gy1 = y_pad.grad.clone()
    r_pad.grad = None; y_pad.grad = None
    ((nr * T64(g["g_nr"], dev)).sum() + (ny * T64(g["g_ny"], dev)).sum()).backward()
    assert relerr(r_pad.grad.cpu(), g["g_r"]) < 1e-10 and relerr(y_pad.grad.cpu(), g["g_y"]) < 1e-10
    assert not torch.equal(gr1, r_pad.grad) or not torch.equal(gy1, y_pad.grad)
    flags.check()


def test_step_vs_golden_fp32(dev):
    """fp32 build vs the reference as shipped (fp32 state, fp64 inner math): states 2e-5 relative to the
    largest entry; gradients 2e-3 (fp32 Jacobian cancellation).  Riemann outcomes may differ only where an
    fp32-rounded quantity sits on a branch threshold (the edge-case fixtures put some there on purpose)."""
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("arz_step_fp32")
    flags = dhts_b200.Flags(dev)
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
    r_pad = t(g["pr"]).requires_grad_(); y_pad = t(g["py"]).requires_grad_()
    nr, ny, nu, case = F.arz_step(r_pad, y_pad, t(g["pu"]), float(g["dx"]), float(g["umax"]), float(g["dt"]), flags,
                                  ueq_pad=t(g["pe"]), want_case=True)
    mism = (case.cpu().numpy() != g["case"])
    assert mism.mean() < 0.02, mism.sum()
    ok = ~(mism[:, :-1] | mism[:, 1:])                      # cells whose two interfaces agree
    a, b = nr.detach().cpu().numpy(), g["nr"]
    assert np.abs(a - b)[ok].max() < 2e-5 * np.abs(b).max()
    a, b = ny.detach().cpu().numpy(), g["ny"]
    assert np.abs(a - b)[ok].max() < 2e-5 * np.abs(b).max()


def test_step_tiles_and_shapes(dev):
    """Lanes longer than one tile, 1-cell lanes, ragged tile tails: tiled kernel == oracle, fwd and bwd."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    for B, N in ((1, 1), (3, 2), (2, 255), (2, 256), (2, 257), (1, 700)):
        dx, umax, dt = 5.0, 30.0, 0.01
        r = rng.uniform(0, 1, (B, N + 2)); u = rng.uniform(0, 1, (B, N + 2)) * umax
        y = O.compute_y(r, u, umax); ue = O.u_eq(r, umax)
        gnr = rng.normal(size=(B, N)); gny = rng.normal(size=(B, N)); gnu = rng.normal(size=(B, N))
        flags = dhts_b200.Flags(dev)
        r_pad = T64(r, dev).requires_grad_(); y_pad = T64(y, dev).requires_grad_()
        nr, ny, nu = F.arz_step(r_pad, y_pad, T64(u, dev), dx, umax, dt, flags)
        ((nr * T64(gnr, dev)).sum() + (ny * T64(gny, dev)).sum() + (nu * T64(gnu, dev)).sum()).backward()
        for b in range(B):
            o = O.arz_step(r[b], y[b], u[b], ue[b], dx, umax, dt)
            assert relerr(nr[b].detach().cpu(), o["nr"]) < 1e-12 and relerr(ny[b].detach().cpu(), o["ny"]) < 1e-12
            dr = np.zeros(N); dy = np.zeros(N)
            for j in range(N):      # fold of g_nu with the true derivative of compute_u
                eps = 1e-5
                if o["nr"][j] >= eps:
                    dr[j] = -o["ny"][j] / o["nr"][j] ** 2 - 0.5 * umax / np.sqrt(o["nr"][j] + eps); dy[j] = 1 / o["nr"][j]
                else:
                    dy[j] = 1 / eps
            gr, gy = O.arz_vjp(o["dqs"], gnr[b] + gnu[b] * dr, gny[b] + gnu[b] * dy)
            assert relerr(r_pad.grad[b].cpu(), gr) < 1e-10 and relerr(y_pad.grad[b].cpu(), gy) < 1e-10
        flags.check()


@pytest.mark.parametrize("ckpt_every", [1, 7, 32, 1000])
def test_rollout_vs_golden_fp64(dev, ckpt_every):
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("arz_rollout_fp64")
    T = int(g["T"])
    flags = dhts_b200.Flags(dev)
    r0 = T64(g["r0"], dev).requires_grad_(); u0 = T64(g["u0"], dev).requires_grad_()
    gr = T64(g["ghost_ru"][:, :, 0], dev).requires_grad_(); gu = T64(g["ghost_ru"][:, :, 1], dev).requires_grad_()
    rT, yT, uT = F.arz_rollout(r0, u0, gr, gu, float(g["dx"]), float(g["umax"]), float(g["dt"]), T,
                               ckpt_every=ckpt_every, flags=flags)
    ((rT * T64(g["w_r"], dev)).sum() + (uT * T64(g["w_u"], dev)).sum()).backward()
    flags.check()
    # north-star tolerance is rtol 1e-5; observed agreement is ~1e-12
    assert relerr(rT.detach().cpu(), g["rT"]) < 1e-9 and relerr(yT.detach().cpu(), g["yT"]) < 1e-9
    assert relerr(uT.detach().cpu(), g["uT"]) < 1e-9
    assert relerr(r0.grad.cpu(), g["g_r0"]) < 1e-8 and relerr(u0.grad.cpu(), g["g_u0"]) < 1e-8
    gg = torch.stack([gr.grad, gu.grad], -1).cpu()
    assert relerr(gg, g["g_ghost"]) < 1e-8


@pytest.mark.parametrize("ckpt_every", [1, 8])
def test_rollout_with_vacuum_vs_live_reference(dev, ckpt_every):
    """The live reference's own rollout over lanes with empty and near-vacuum stretches and vacuum ghosts (fixture
    arz_rollout_vac_fp64, oracle/gen_golden_vac.py): 128-cell lanes, so ckpt_every = 1 runs the staged forward kernel with
    stored outcomes and the ring adjoint (four cells per thread, vacuum votes), ckpt_every = 8 the recompute path."""
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("arz_rollout_vac_fp64")
    T = int(g["T"])
    flags = dhts_b200.Flags(dev)
    r0 = T64(g["r0"], dev).requires_grad_(); u0 = T64(g["u0"], dev).requires_grad_()
    gr = T64(g["ghost_ru"][:, :, 0], dev).requires_grad_(); gu = T64(g["ghost_ru"][:, :, 1], dev).requires_grad_()
    rT, yT, uT = F.arz_rollout(r0, u0, gr, gu, float(g["dx"]), float(g["umax"]), float(g["dt"]), T,
                               ckpt_every=ckpt_every, flags=flags)
    ((rT * T64(g["w_r"], dev)).sum() + (uT * T64(g["w_u"], dev)).sum()).backward()
    flags.check()
    assert (g["rT"] < 1e-5).any()
    assert relerr(rT.detach().cpu(), g["rT"]) < 1e-9 and relerr(yT.detach().cpu(), g["yT"]) < 1e-9
    assert relerr(uT.detach().cpu(), g["uT"]) < 1e-9
    assert relerr(r0.grad.cpu(), g["g_r0"]) < 1e-8 and relerr(u0.grad.cpu(), g["g_u0"]) < 1e-8
    assert relerr(torch.stack([gr.grad, gu.grad], -1).cpu(), g["g_ghost"]) < 1e-8


def test_rollout_vs_golden_fp32(dev):
    """fp32 build vs the reference as shipped, T=300: states 1e-4, gradients 2e-3 of the largest entry."""
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("arz_rollout_fp32")
    T = int(g["T"])
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
    flags = dhts_b200.Flags(dev)
    r0 = t(g["r0"]).requires_grad_(); u0 = t(g["u0"]).requires_grad_()
    rT, yT, uT = F.arz_rollout(r0, u0, t(g["ghost_ru"][:, :, 0]), t(g["ghost_ru"][:, :, 1]), float(g["dx"]),
                               float(g["umax"]), float(g["dt"]), T, ckpt_every=25, flags=flags)
    ((rT * t(g["w_r"])).sum() + (uT * t(g["w_u"])).sum()).backward()
    flags.check()
    assert relerr(rT.detach().cpu(), g["rT"]) < 1e-4 and relerr(uT.detach().cpu(), g["uT"]) < 1e-4
    assert relerr(r0.grad.cpu(), g["g_r0"]) < 2e-3 and relerr(u0.grad.cpu(), g["g_u0"]) < 2e-3


@pytest.mark.parametrize("B,N,T,K", [(1, 10, 500, 32), (37, 10, 60, 8), (5, 100, 50, 16), (3, 1024, 24, 8),
                                      (300, 33, 20, 5), (3, 1024, 24, 1), (301, 64, 20, 1), (5, 128, 3, 1),
                                      (700, 1024, 6, 1), (2, 4096, 5, 1)])
def test_rollout_vs_oracle_shapes(dev, B, N, T, K):
    """C1's shape (1 x 10 x 500), many short lanes per CTA, C5's lane shape (1024 cells), ragged groups."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(B * 1000 + N)
    dx = rng.uniform(4.0, 6.0, B); umax = rng.uniform(25.0, 35.0, B); dt = 0.01
    r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 1, (B, N)) * umax[:, None]
    gh = np.stack([rng.uniform(0, 1, (B, 2)), rng.uniform(0, 1, (B, 2)) * umax[:, None]], -1)
    wr = rng.normal(size=(B, N)); wu = rng.normal(size=(B, N)) / 30; wy = rng.normal(size=(B, N)) / 30
    flags = dhts_b200.Flags(dev)
    tr = T64(r0, dev).requires_grad_(); tu = T64(u0, dev).requires_grad_()
    tgr = T64(gh[:, :, 0], dev).requires_grad_(); tgu = T64(gh[:, :, 1], dev).requires_grad_()
    rT, yT, uT = F.arz_rollout(tr, tu, tgr, tgu, T64(dx, dev), T64(umax, dev), dt, T, ckpt_every=K, flags=flags)
    ((rT * T64(wr, dev)).sum() + (uT * T64(wu, dev)).sum() + (yT * T64(wy, dev)).sum()).backward()
    flags.check()
    o = O.arz_rollout(r0, u0, gh, dx, umax, dt, T, g_rT=wr, g_yT=wy, g_uT=wu)
    assert o["cfl"] == 0
    assert relerr(rT.detach().cpu(), o["rT"]) < 1e-9 and relerr(uT.detach().cpu(), o["uT"]) < 1e-9
    assert relerr(tr.grad.cpu(), o["g_r0"]) < 1e-8 and relerr(tu.grad.cpu(), o["g_u0"]) < 1e-8
    assert relerr(torch.stack([tgr.grad, tgu.grad], -1).cpu(), o["g_ghost"]) < 1e-8


@pytest.mark.parametrize("B,N,T,K", [(6, 1024, 40, 1), (6, 1024, 40, 8), (37, 64, 30, 1), (9, 128, 25, 1), (4, 100, 30, 5)])
def test_rollout_with_vacuum_vs_oracle(dev, B, N, T, K):
    """Lanes with EMPTY stretches (r = 0), near-vacuum cells on both sides of eps = 1e-5 and vacuum ghosts: the warp-voted
    vacuum variants of the forward / adjoint sweeps (fix-ups of w, f00, f11; clamps at eps), on the staged + stored-outcome
    path (ckpt_every = 1: the adjoint picks its variant from its own cells, the outcomes come from the forward pass), the
    recompute path and the many-lanes-per-CTA shapes.  _arz.py:225-322 (vacuum branches of the case tree), darz.py:217-233."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(7 * B + N + K)
    dx = rng.uniform(4.0, 6.0, B); umax = rng.uniform(25.0, 35.0, B); dt = 0.01
    r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 1, (B, N)) * umax[:, None]
    for b in range(B):
        for _ in range(3):      # empty stretches and near-vacuum stretches (some cells just below, some just above eps)
            a = int(rng.integers(0, N - 8)); w = int(rng.integers(3, max(4, N // 6)))
            r0[b, a:a + w] = 0.0
            a = int(rng.integers(0, N - 8)); w = int(rng.integers(3, max(4, N // 8)))
            r0[b, a:a + w] = 10.0 ** rng.uniform(-7, -4, size=r0[b, a:a + w].shape)
    gh = np.stack([rng.uniform(0, 1, (B, 2)), rng.uniform(0, 1, (B, 2)) * umax[:, None]], -1)
    gh[::2, 0, 0] = 0.0; gh[1::3, 1, 0] = 3e-6                  # vacuum ghosts
    wr = rng.normal(size=(B, N)); wu = rng.normal(size=(B, N)) / 30
    flags = dhts_b200.Flags(dev)
    tr = T64(r0, dev).requires_grad_(); tu = T64(u0, dev).requires_grad_()
    tgr = T64(gh[:, :, 0], dev).requires_grad_(); tgu = T64(gh[:, :, 1], dev).requires_grad_()
    rT, yT, uT = F.arz_rollout(tr, tu, tgr, tgu, T64(dx, dev), T64(umax, dev), dt, T, ckpt_every=K, flags=flags)
    ((rT * T64(wr, dev)).sum() + (uT * T64(wu, dev)).sum()).backward()
    flags.check()
    o = O.arz_rollout(r0, u0, gh, dx, umax, dt, T, g_rT=wr, g_uT=wu)
    assert o["cfl"] == 0
    assert (o["rT"] < 1e-5).any()                                # vacuum cells survive to the end of the rollout
    assert relerr(rT.detach().cpu(), o["rT"]) < 1e-9 and relerr(uT.detach().cpu(), o["uT"]) < 1e-9
    assert relerr(tr.grad.cpu(), o["g_r0"]) < 1e-8 and relerr(tu.grad.cpu(), o["g_u0"]) < 1e-8
    assert relerr(torch.stack([tgr.grad, tgu.grad], -1).cpu(), o["g_ghost"]) < 1e-8


@pytest.mark.parametrize("B,N,T,K", [(9, 1024, 14, 1), (300, 1024, 5, 1), (9, 1024, 14, 8), (37, 10, 30, 8), (21, 128, 9, 1)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_rollout_uniform_geometry_equals_per_lane_arrays(dev, dtype, B, N, T, K):
    """dx / umax given as two scalars (include/dhts.h: dx = umax = NULL, dx_all / umax_all -- every lane of the reference's
    drivers has the same cell length and speed limit) select kernels that read the lane constants from their parameter
    block; the arithmetic is the same, so states and gradients equal the per-lane-array call BITWISE."""
    import dhts_b200
    from dhts_b200 import functional as F
    g = torch.Generator(device=dev).manual_seed(B * 31 + N)
    rnd = lambda *s: torch.rand(s, generator=g, dtype=dtype, device=dev)
    r0, u0, gr, gu = rnd(B, N), rnd(B, N) * 30, rnd(B, 2), rnd(B, 2) * 30
    w = torch.randn((B, N), generator=g, dtype=dtype, device=dev)
    flags = dhts_b200.Flags(dev)
    outs = []
    for dx, um in ((5.0, 30.0), (torch.full((B,), 5.0, dtype=dtype, device=dev), torch.full((B,), 30.0, dtype=dtype, device=dev))):
        a = r0.clone().requires_grad_(); b = u0.clone().requires_grad_()
        c = gr.clone().requires_grad_(); d = gu.clone().requires_grad_()
        rT, yT, uT = F.arz_rollout(a, b, c, d, dx, um, 0.01, T, ckpt_every=K, flags=flags)
        ((rT * w).sum() + (uT * w).sum() / 30).backward()
        outs.append((rT.detach(), yT.detach(), uT.detach(), a.grad, b.grad, c.grad, d.grad))
    flags.check()
    for x, y in zip(*outs):
        assert torch.equal(x, y)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_rollout_tma_paths_equal_plain_paths(dev, dtype):
    """Every state stored: the forward's staged TMA bulk stores (shared-memory staging ring, cp.async.bulk to HBM) and
    the adjoint's TMA state ring (cp.async.bulk into shared memory, mbarrier completion) against the per-thread
    store / register-prefetch variants.  Same arithmetic, so gradients must agree to rounding (observed: bitwise
    or 1 ulp) -- persistent loop over more lane groups than CTAs, ragged last group, fewer steps than ring
    stages, stored-speed first step."""
    import os
    import dhts_b200
    from dhts_b200 import functional as F
    rng = np.random.default_rng(5)
    variants = {"tma": {}, "ring2": {"DHTS_ARZ_RING": "2"}, "plain_adj": {"DHTS_ARZ_RING": "0"},
                "plain": {"DHTS_ARZ_RING": "0", "DHTS_ARZ_STAGE": "0"}}
    for B, N, T in ((701, 1024, 7), (301, 64, 9), (9, 256, 2), (40, 2048, 5)):
        r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 30, (B, N))
        gr = rng.uniform(0, 1, (B, 2)); gu = rng.uniform(0, 30, (B, 2)); w = rng.normal(size=(B, N))
        t = lambda a: torch.tensor(a, dtype=dtype, device=dev)
        out = {}
        for name, env in variants.items():
            os.environ.update(env)
            try:
                flags = dhts_b200.Flags(dev)
                tr, tu = t(r0).requires_grad_(), t(u0).requires_grad_()
                tgr, tgu = t(gr).requires_grad_(), t(gu).requires_grad_()
                rT, yT, uT = F.arz_rollout(tr, tu, tgr, tgu, 5.0, 30.0, 0.01, T, ckpt_every=1, flags=flags)
                ((rT * t(w)).sum() + (uT * t(w / 30)).sum()).backward()
                flags.check()
                out[name] = [x.clone() for x in (rT.detach(), uT.detach(), tr.grad, tu.grad, tgr.grad, tgu.grad)]
            finally:
                for k in env:
                    del os.environ[k]
        tol = 1e-13 if dtype == torch.float64 else 1e-5       # same arithmetic; the compiler may contract differently
        for name in ("tma", "ring2", "plain_adj"):
            for a, b in zip(out[name], out["plain"]):
                assert relerr(a.cpu(), b.cpu()) < tol, (B, N, T, name)


def test_rollout_equals_chained_steps(dev):
    """Fused rollout == T chained single-step operators (the per-step contract) up to FMA-contraction
    differences between the two kernels (1e-12)."""
    import dhts_b200
    from dhts_b200 import functional as F
    rng = np.random.default_rng(11)
    B, N, T, dx, umax, dt = 4, 50, 30, 5.0, 30.0, 0.01
    r0 = T64(rng.uniform(0, 1, (B, N)), dev); u0 = T64(rng.uniform(0, 30, (B, N)), dev)
    gr = T64(rng.uniform(0, 1, (B, 2)), dev); gu = T64(rng.uniform(0, 30, (B, 2)), dev)
    flags = dhts_b200.Flags(dev)
    rT, yT, uT = F.arz_rollout(r0, u0, gr, gu, dx, umax, dt, T, flags=flags)
    gy = F.compute_y(gr, gu, umax)
    r, y, u = r0, F.compute_y(r0, u0, umax), u0
    for _ in range(T):
        cat = lambda a, b: torch.cat([a[:, :1], b, a[:, 1:]], 1)
        r, y, u = F.arz_step(cat(gr, r), cat(gy, y), cat(gu, u), dx, umax, dt, flags)
    for a, b in ((r, rT), (y, yT), (u, uT)):
        assert relerr(a.cpu(), b.cpu()) < 1e-12


def test_cfl_flag_maps_to_assertion(dev):
    import dhts_b200
    from dhts_b200 import functional as F
    r0 = torch.full((1, 8), 0.5, dtype=torch.float64, device=dev); u0 = torch.full((1, 8), 29.0, dtype=torch.float64, device=dev)
    g = torch.full((1, 2), 0.5, dtype=torch.float64, device=dev)
    flags = dhts_b200.Flags(dev)
    F.arz_rollout(r0, u0, g, g * 58, 0.1, 30.0, 0.01, 1, flags=flags)     # dx = 0.1 violates CFL
    with pytest.raises(AssertionError, match="CFL"):
        flags.check()


def test_long_lane_falls_back_to_step_kernels(dev):
    """A lane that does not fit in shared memory runs through the tiled per-step CUDA kernels."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    B, N, T = 1, 6000, 3
    r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 30, (B, N)); gh = np.array([[[0.3, 10.0], [0.6, 20.0]]])
    w = rng.normal(size=(B, N))
    flags = dhts_b200.Flags(dev)
    tr = T64(r0, dev).requires_grad_(); tu = T64(u0, dev)
    rT, yT, uT = F.arz_rollout(tr, tu, T64(gh[:, :, 0], dev), T64(gh[:, :, 1], dev), 5.0, 30.0, 0.01, T, flags=flags)
    (rT * T64(w, dev)).sum().backward()
    o = O.arz_rollout(r0, u0, gh, 5.0, 30.0, 0.01, T, g_rT=w)
    assert relerr(rT.detach().cpu(), o["rT"]) < 1e-12 and relerr(tr.grad.cpu(), o["g_r0"]) < 1e-9


def test_rollout_cfl_slow_path_and_arena_and_unaligned(dev):
    """(a) speeds above V/4 but below V = dx/dt: the sufficient CFL tests fail, the exact per-branch test must pass
    (no flag) and the states stay right; (b) checkpoints written into a caller-owned arena; (c) row starts that are
    not 16-byte aligned take the tiled per-step kernels -- all three against the oracle."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    B, N, T, dx, umax, dt = 6, 64, 40, 1.0, 30.0, 0.01          # V = 100, V/4 = 25 < typical u
    r0 = rng.uniform(0.05, 0.9, (B, N)); u0 = rng.uniform(20.0, 29.0, (B, N))
    gh = np.stack([rng.uniform(0.1, 0.9, (B, 2)), rng.uniform(20, 29, (B, 2))], -1)
    w = rng.normal(size=(B, N))
    o = O.arz_rollout(r0, u0, gh, dx, umax, dt, T, g_rT=w, g_uT=w / umax)
    assert o["cfl"] == 0
    n_ck = F.arz_ckpt_elems(B, N, T, 1, torch.float64)      # the T states and, where stored, the interface outcomes
    assert n_ck >= T * 2 * B * N
    arena = torch.empty(n_ck + 5, dtype=torch.float64, device=dev)
    for mode in ("plain", "arena", "unaligned"):
        flags = dhts_b200.Flags(dev)
        if mode == "unaligned":
            big = torch.zeros(B * N + 1, dtype=torch.float64, device=dev)
            big[1:] = T64(r0, dev).reshape(-1)
            tr = big[1:].view(B, N).detach().requires_grad_()     # data_ptr is 8 mod 16
            assert tr.data_ptr() % 16 == 8
        else:
            tr = T64(r0, dev).requires_grad_()
        tu = T64(u0, dev).requires_grad_()
        rT, yT, uT = F.arz_rollout(tr, tu, T64(gh[:, :, 0], dev), T64(gh[:, :, 1], dev), dx, umax, dt, T, ckpt_every=1,
                                   flags=flags, ckpt_buffer=arena if mode == "arena" else None)
        ((rT * T64(w, dev)).sum() + (uT * T64(w / umax, dev)).sum()).backward()
        assert flags.check()[0] == 0
        assert relerr(rT.detach().cpu(), o["rT"]) < 1e-11 and relerr(uT.detach().cpu(), o["uT"]) < 1e-11, mode
        assert relerr(tr.grad.cpu(), o["g_r0"]) < 1e-9 and relerr(tu.grad.cpu(), o["g_u0"]) < 1e-9, mode
    with pytest.raises(ValueError, match="ckpt_buffer"):
        F.arz_rollout(T64(r0, dev).requires_grad_(), T64(u0, dev), T64(gh[:, :, 0], dev), T64(gh[:, :, 1], dev), dx, umax,
                      dt, T, ckpt_every=1, ckpt_buffer=arena[:10])


def test_rollout_plan_fits_memory(dev):
    from dhts_b200 import functional as F
    chunk, K = F.arz_rollout_plan(65536, 1024, 1000, torch.float64, dev)
    free, _ = torch.cuda.mem_get_info(dev)
    assert K == 1 and 296 <= chunk <= 65536 and chunk * 1024 * 1000 * 2 * 8 <= 0.6 * free + 1
    nchunk = -(-65536 // chunk)
    assert (nchunk - 1) * chunk < 65536 <= nchunk * chunk
    assert F.arz_rollout_plan(100, 64, 10, torch.float64, dev) == (100, 1)
    chunk, K = F.arz_rollout_plan(4096, 1024, 10 ** 7, torch.float64, dev)        # nothing fits: sparse checkpoints
    assert K == 32 and chunk >= 1


@pytest.mark.parametrize("ckpt_every,N", [(1, 1024), (16, 1024), (1, 40)])
def test_nan_gradient_flag_maps_to_assertion(dev, ckpt_every, N):
    """A NaN that enters the adjoint (here: through the terminal adjoint of ONE cell) stays in that cell through every step,
    so the single test of the final adjoint raises the reference's NaN assert (dmacro_lane.py:308); without a NaN the flag
    stays clear.  Covers the TMA-ring, the recompute and the small-lane paths of the adjoint kernel."""
    from dhts_b200 import Flags, functional as F
    from dhts_b200._lib import FLAG_NAN_GRAD
    rng = np.random.default_rng(5)
    B, T = 3, 40
    r0 = T64(rng.uniform(0.1, 0.9, (B, N)), dev); u0 = T64(rng.uniform(0, 30, (B, N)), dev)
    gr = T64(rng.uniform(0.1, 0.9, (B, 2)), dev); gu = T64(rng.uniform(0, 30, (B, 2)), dev)
    for poison in (False, True):
        tr = r0.clone().requires_grad_()
        flags = Flags(dev)
        rT, yT, uT = F.arz_rollout(tr, u0, gr, gu, 5.0, 30.0, 0.01, T, ckpt_every=ckpt_every, flags=flags)
        w = torch.ones_like(rT)
        if poison:
            w[1, N // 2] = float("nan")
        (rT * w).sum().backward()
        bits, _ = flags.read()
        assert bool(bits & FLAG_NAN_GRAD) == poison
        assert bool(torch.isnan(tr.grad[1]).any()) == poison and not bool(torch.isnan(tr.grad[0]).any())
        if poison:
            with pytest.raises(AssertionError):
                flags.check()
        else:
            flags.check()


def test_rollout_per_step_coupling_vs_live_reference(dev):
    """Per-step ghosts + a loss on the state before every step, against the live fp64 reference (arz_perstep_fp64.npz):
    history, final state, gradients wrt the initial state and wrt every step's ghost cells."""
    import dhts_b200
    from dhts_b200 import functional as F
    g = golden("arz_perstep_fp64")
    T, umax = int(g["T"]), float(g["umax"])
    flags = dhts_b200.Flags(dev)
    r0 = T64(g["r0"], dev).requires_grad_(); u0 = T64(g["u0"], dev).requires_grad_()
    gr = T64(g["ghost_ru"][..., 0], dev).requires_grad_(); gu = T64(g["ghost_ru"][..., 1], dev).requires_grad_()
    rT, yT, uT, rh, yh = F.arz_rollout(r0, u0, gr, gu, float(g["dx"]), umax, float(g["dt"]), T, ckpt_every=1, flags=flags,
                                       return_history=True)
    uh = yh / rh.clamp(min=1e-5) + F.u_eq(rh.clamp(min=1e-5), umax)            # compute_u, model/macro/_arz.py:126-131
    loss = (rh * T64(g["w_r"], dev)).sum() + (uh * T64(g["w_u"], dev)).sum() + (rT * T64(g["wT_r"], dev)).sum() \
        + (uT * T64(g["wT_u"], dev)).sum()
    loss.backward()
    flags.check()
    assert relerr(rh.detach().cpu(), g["r_hist"]) < 1e-9 and relerr(uh.detach().cpu(), g["u_hist"]) < 1e-9
    assert relerr(rT.detach().cpu(), g["rT"]) < 1e-9 and relerr(uT.detach().cpu(), g["uT"]) < 1e-9
    assert abs(float(loss) - float(g["loss"].sum())) < 1e-9 * abs(float(g["loss"].sum())) + 1e-9
    assert relerr(r0.grad.cpu(), g["g_r0"]) < 1e-8 and relerr(u0.grad.cpu(), g["g_u0"]) < 1e-8
    assert relerr(torch.stack([gr.grad, gu.grad], -1).cpu(), g["g_ghost"]) < 1e-8


@pytest.mark.parametrize("B,N,T", [(5, 1024, 20), (37, 10, 40), (300, 33, 12)])
def test_rollout_per_step_coupling_vs_oracle_shapes(dev, B, N, T):
    """Same, on the bench's lane shape (4 cells per thread), many short lanes per CTA and ragged groups, vs the oracle."""
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    rng = np.random.default_rng(B + N)
    umax, dx, dt = 30.0, 5.0, 0.01
    r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 1, (B, N)) * umax
    gh = np.stack([rng.uniform(0, 1, (T, B, 2)), rng.uniform(0, 1, (T, B, 2)) * umax], -1)
    g_hist = rng.normal(size=(T, B, N, 2)) * np.array([1.0, 1.0 / umax])
    wr = rng.normal(size=(B, N)); wu = rng.normal(size=(B, N)) / umax
    flags = dhts_b200.Flags(dev)
    tr = T64(r0, dev).requires_grad_(); tu = T64(u0, dev).requires_grad_()
    tgr = T64(gh[..., 0], dev).requires_grad_(); tgu = T64(gh[..., 1], dev).requires_grad_()
    rT, yT, uT, rh, yh = F.arz_rollout(tr, tu, tgr, tgu, dx, umax, dt, T, ckpt_every=1, flags=flags, return_history=True)
    ((rh * T64(g_hist[..., 0], dev)).sum() + (yh * T64(g_hist[..., 1], dev)).sum() + (rT * T64(wr, dev)).sum()
     + (uT * T64(wu, dev)).sum()).backward()
    flags.check()
    o = O.arz_rollout(r0, u0, gh, dx, umax, dt, T, g_rT=wr, g_uT=wu, g_hist=g_hist)
    assert o["cfl"] == 0
    assert relerr(rT.detach().cpu(), o["rT"]) < 1e-9 and relerr(uT.detach().cpu(), o["uT"]) < 1e-9
    assert relerr(tr.grad.cpu(), o["g_r0"]) < 1e-8 and relerr(tu.grad.cpu(), o["g_u0"]) < 1e-8
    assert relerr(torch.stack([tgr.grad, tgu.grad], -1).cpu(), o["g_ghost"]) < 1e-8
