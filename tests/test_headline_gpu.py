"""The HEADLINE configuration against the oracle: lanes of BASELINE.json configs[4]'s shape (1024 cells / 64 vehicles)
rolled out for the full T = 1000 steps forward + adjoint in the modes bench.py runs -- ARZ with every state stored
(ckpt_every = 1: TMA staging ring forward, TMA ring adjoint, 250 ring wrap-arounds) through ONE checkpoint arena
reused by consecutive lane chunks, and with sparse checkpoints + segment recompute (ckpt_every = 32); IDM with
ckpt_every = 16 -- with the bench's inputs and loss (example/inverse/macro.py:226-241, micro.py:221-236;
loop semantics of example/inverse/_inverse.py:91-99).  bench.py repeats the same check on sampled lanes of every timed
run (`parity_check`).  fp64: states <= 1e-9, gradients <= 1e-8 of the largest entry (north star: rtol 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu

SEED = 20221008


def test_arz_headline_modes_vs_oracle(dev):
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    import os
    O.set_threads(len(os.sched_getaffinity(0)))
    B, N, T, dx, umax, dt = 704, 1024, 1000, 5.0, 30.0, 0.01
    g = torch.Generator().manual_seed(SEED)
    rnd = lambda *s: torch.rand(s, generator=g, dtype=torch.float32).to(torch.float64)
    r0, u0, gr, gu, tr, tu = rnd(B, N), rnd(B, N) * umax, rnd(B, 2), rnd(B, 2) * umax, rnd(B, N), rnd(B, N) * umax
    f = O.arz_rollout(r0.numpy(), u0.numpy(), np.stack([gr.numpy(), gu.numpy()], -1), dx, umax, dt, T)
    o = O.arz_rollout(r0.numpy(), u0.numpy(), np.stack([gr.numpy(), gu.numpy()], -1), dx, umax, dt, T,
                      g_rT=2.0 * (f["rT"] - tr.numpy()), g_uT=2.0 * (f["uT"] - tu.numpy()))
    assert o["cfl"] == 0
    d = {k: v.to(dev) for k, v in dict(r0=r0, u0=u0, gr=gr, gu=gu, tr=tr, tu=tu).items()}
    flags = dhts_b200.Flags(dev)
    chunk = B // 2
    arena = torch.empty(F.arz_ckpt_elems(chunk, N, T, 1, torch.float64), dtype=torch.float64, device=dev)     # one arena (states + interface outcomes), reused by both chunks

    def run(K, chunks, arena):
        out = {k: torch.empty((B, N), dtype=torch.float64, device=dev) for k in ("rT", "uT", "g_r0", "g_u0")}
        loss = 0.0
        for lo, hi in chunks:
            a = d["r0"][lo:hi].detach().requires_grad_(); b = d["u0"][lo:hi].detach().requires_grad_()
            rT, yT, uT = F.arz_rollout(a, b, d["gr"][lo:hi], d["gu"][lo:hi], dx, umax, dt, T, ckpt_every=K, flags=flags,
                                       ckpt_buffer=arena)
            l = ((rT - d["tr"][lo:hi]) ** 2).sum() + ((uT - d["tu"][lo:hi]) ** 2).sum()
            l.backward()
            out["rT"][lo:hi] = rT.detach(); out["uT"][lo:hi] = uT.detach(); out["g_r0"][lo:hi] = a.grad; out["g_u0"][lo:hi] = b.grad
            loss += float(l)
        flags.check()
        return out, loss

    k1, loss1 = run(1, [(0, chunk), (chunk, B)], arena)
    for name, tol in (("rT", 1e-9), ("uT", 1e-9), ("g_r0", 1e-8), ("g_u0", 1e-8)):
        assert relerr(k1[name].cpu(), o[name]) < tol, name
    ref_loss = float(((f["rT"] - tr.numpy()) ** 2).sum() + ((f["uT"] - tu.numpy()) ** 2).sum())
    assert abs(loss1 - ref_loss) <= 1e-9 * abs(ref_loss)
    del arena
    k32, loss32 = run(32, [(0, B)], None)
    for name, tol in (("rT", 1e-9), ("uT", 1e-9), ("g_r0", 1e-8), ("g_u0", 1e-8)):
        assert relerr(k32[name].cpu(), o[name]) < tol, name
    # same forward kernel -> identical states; the recompute adjoint is another instantiation -> gradients to rounding
    assert torch.equal(k1["rT"], k32["rT"]) and torch.equal(k1["uT"], k32["uT"])
    for name in ("g_r0", "g_u0"):
        assert float((k1[name] - k32[name]).abs().max()) <= 1e-11 * float(k1[name].abs().max()), name
    assert loss32 == pytest.approx(loss1, rel=1e-12)


def test_idm_headline_mode_vs_oracle(dev):
    import dhts_b200
    from dhts_b200 import functional as F
    from oracle import oracle as O
    L, n, T, umax, dt = 1024, 64, 1000, 30.0, 0.01
    V = L * n
    g = torch.Generator().manual_seed(SEED + 7919)
    rnd = lambda *s: torch.rand(s, generator=g, dtype=torch.float32).to(torch.float64)
    p0 = (torch.arange(n, dtype=torch.float64)[None, :] * 20.0 + rnd(L, n) * 10.0).reshape(V)
    v0 = 9.0 + 12.0 * rnd(V)
    par = torch.stack([(1.5 + 0.5 * rnd(V)) * umax, (1.0 + 0.5 * rnd(V)) * umax, (0.8 + 0.4 * rnd(V)) * umax, 1.0 + rnd(V),
                       0.2 + 0.4 * rnd(V), torch.full((V,), 5.0, dtype=torch.float64)])
    tp = p0 + dt * T * 15.0 + rnd(V); tv = 9.0 + 12.0 * rnd(V)
    off = (torch.arange(L + 1) * n).to(torch.int32)
    head = torch.tensor([[1000.0, 0.0]], dtype=torch.float64).repeat(L, 1)
    f = O.idm_rollout(p0.numpy(), v0.numpy(), par.numpy(), off.numpy(), head.numpy(), dt, T)
    o = O.idm_rollout(p0.numpy(), v0.numpy(), par.numpy(), off.numpy(), head.numpy(), dt, T,
                      g_pT=2.0 * (f["pT"] - tp.numpy()), g_vT=2.0 * (f["vT"] - tv.numpy()))
    flags = dhts_b200.Flags(dev)
    a = p0.to(dev).requires_grad_(); b = v0.to(dev).requires_grad_()
    pT, vT = F.idm_rollout(a, b, par.to(dev), off.to(dev), head.to(dev), dt, T, ckpt_every=16, flags=flags, max_lane=n)
    (((pT - tp.to(dev)) ** 2).sum() + ((vT - tv.to(dev)) ** 2).sum()).backward()
    bits, ncol = flags.check()
    assert ncol == o["ncol"] == 0
    assert relerr(pT.detach().cpu(), o["pT"]) < 1e-9 and relerr(vT.detach().cpu(), o["vT"]) < 1e-9
    assert relerr(a.grad.cpu(), o["g_p0"]) < 1e-8 and relerr(b.grad.cpu(), o["g_v0"]) < 1e-8
