"""Size-independent properties at BASELINE.json's full per-GPU shapes (65536 lanes x 1024 cells; 4,194,304 vehicles
in 65536 lanes), where the oracle cannot follow: lanes are independent, so results must be BITWISE invariant under
sharding, lane permutation and the checkpoint interval, and a lane with zero loss weight must get an exactly zero
gradient.  (Steps are kept at 64 / 96 so the test stays in seconds; the time loop is the same code for any T.)
The base run uses ckpt_every = 1 -- every state stored, TMA staging ring forward, TMA ring adjoint: the mode
bench.py times (69 GB of stored states for the 65536 lanes x 64 steps here); the oracle comparison of that mode at
the full T = 1000 is tests/test_headline_gpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SEED = 20221008


def _arz_pass(F, flags, r0, u0, gr, gu, w, T, K):
    r0 = r0.detach().requires_grad_(); u0 = u0.detach().requires_grad_()
    rT, yT, uT = F.arz_rollout(r0, u0, gr, gu, 5.0, 30.0, 0.01, T, ckpt_every=K, flags=flags)
    ((rT * w).sum() + (uT * w).sum() / 30.0).backward()
    return rT.detach(), uT.detach(), r0.grad, u0.grad


def test_arz_full_batch_invariances(dev):
    import dhts_b200
    from dhts_b200 import functional as F
    B, N, T = 65536, 1024, 64
    g = torch.Generator(device=dev).manual_seed(SEED)
    rnd = lambda *s: torch.rand(s, generator=g, dtype=torch.float64, device=dev)
    r0, u0, gr, gu = rnd(B, N), rnd(B, N) * 30, rnd(B, 2), rnd(B, 2) * 30
    w = torch.randn((B, N), generator=g, dtype=torch.float64, device=dev)
    w[1::2] = 0.0                                            # odd lanes carry no loss
    flags = dhts_b200.Flags(dev)
    full = _arz_pass(F, flags, r0, u0, gr, gu, w, T, 1)
    flags.check()
    assert all(torch.isfinite(x).all() for x in full)
    assert (full[2][1::2] == 0).all() and (full[3][1::2] == 0).all() and full[2][0::2].abs().max() > 0
    # shard == unshard, bitwise (three uneven shards)
    for lo, hi in ((0, 20000), (20000, 20001), (20001, B)):
        torch.cuda.empty_cache()          # the 69 GB state arena of the previous pass goes back to the driver (no fragmentation)
        part = _arz_pass(F, flags, r0[lo:hi], u0[lo:hi], gr[lo:hi], gu[lo:hi], w[lo:hi], T, 1)
        for a, b in zip(full, part):
            assert torch.equal(a[lo:hi], b)
    # lane permutation equivariance, bitwise
    perm = torch.randperm(B, generator=g, device=dev)
    torch.cuda.empty_cache()
    pp = _arz_pass(F, flags, r0[perm], u0[perm], gr[perm], gu[perm], w[perm], T, 1)
    for a, b in zip(full, pp):
        assert torch.equal(a[perm], b)
    del pp
    torch.cuda.empty_cache()
    # sparse checkpoints + segment recompute is a separate instantiation of the adjoint kernel (the compiler is free to
    # contract its multiply-adds differently): same forward states bitwise, gradients to rounding
    k32 = _arz_pass(F, flags, r0, u0, gr, gu, w, T, 32)
    assert torch.equal(full[0], k32[0]) and torch.equal(full[1], k32[1])
    for a, b in zip(full[2:], k32[2:]):
        assert float((a - b).abs().max()) <= 1e-11 * float(a.abs().max())
    # within that mode the checkpoint interval changes what is stored, not what is computed
    for K in (7, 64, 200):
        kk = _arz_pass(F, flags, r0, u0, gr, gu, w, T, K)
        for a, b in zip(k32, kk):
            assert torch.equal(a, b)
    flags.check()


def test_idm_full_batch_invariances(dev):
    import dhts_b200
    from dhts_b200 import functional as F
    L, n, T = 65536, 64, 96
    V = L * n
    g = torch.Generator(device=dev).manual_seed(SEED + 1)
    rnd = lambda *s: torch.rand(s, generator=g, dtype=torch.float64, device=dev)
    p0 = (torch.arange(n, dtype=torch.float64, device=dev)[None] * 20.0 + rnd(L, n) * 10).reshape(V)
    v0 = 9.0 + 12.0 * rnd(V)
    par = torch.stack([(1.5 + 0.5 * rnd(V)) * 30, (1.0 + 0.5 * rnd(V)) * 30, (0.8 + 0.4 * rnd(V)) * 30, 1.0 + rnd(V),
                       0.2 + 0.4 * rnd(V), torch.full((V,), 5.0, dtype=torch.float64, device=dev)])
    off = (torch.arange(L + 1, device=dev) * n).to(torch.int32)
    head = torch.tensor([[1000.0, 0.0]], dtype=torch.float64, device=dev).repeat(L, 1)
    w = torch.randn(V, generator=g, dtype=torch.float64, device=dev)
    w.view(L, n)[1::2] = 0.0
    flags = dhts_b200.Flags(dev)

    def run(p0, v0, par, off, head, w, K):
        p0 = p0.detach().requires_grad_(); v0 = v0.detach().requires_grad_(); head = head.detach().requires_grad_()
        pT, vT = F.idm_rollout(p0, v0, par, off, head, 0.01, T, ckpt_every=K, flags=flags, max_lane=n)
        ((pT * w).sum() + (vT * w).sum()).backward()
        return pT.detach(), vT.detach(), p0.grad, v0.grad, head.grad

    full = run(p0, v0, par, off, head, w, 32)
    bits, ncol = flags.check()
    assert ncol == 0 and all(torch.isfinite(x).all() for x in full)
    assert (full[2].view(L, n)[1::2] == 0).all() and (full[4][1::2] == 0).all() and full[2].abs().max() > 0
    for lo, hi in ((0, 30001), (30001, L)):
        a, b = lo * n, hi * n
        part = run(p0[a:b], v0[a:b], par[:, a:b].contiguous(), (off[lo:hi + 1] - off[lo]).to(torch.int32), head[lo:hi],
                   w[a:b], 32)
        for x, y in zip(full[:4], part[:4]):
            assert torch.equal(x[a:b], y)
        assert torch.equal(full[4][lo:hi], part[4])
    for K in (5, 16):
        kk = run(p0, v0, par, off, head, w, K)
        for x, y in zip(full, kk):
            assert torch.equal(x, y)
    flags.check()
