"""macro<->micro exchange kernels (dhts_m2c_*, dhts_c2m_*) through the C ABI against the plain-torch restatement of
road/network/conversion.py:15-171 in tests/cpu_standin.py, batched over junctions, values and gradients.
(The exchange inside a running network is covered by the hybrid-chain fixtures in tests/test_dropin_gpu.py.)"""
import numpy as np
import pytest
import torch

import cpu_standin as S
from conftest import relerr

pytestmark = pytest.mark.gpu


def _leafs(arrs, dev=None):
    return [torch.tensor(a, dtype=torch.float64, device=dev).requires_grad_() for a in arrs]


def test_macro_to_micro_batched(dev):
    from dhts_b200 import functional as F
    rng = np.random.default_rng(7)
    J, dt = 513, 0.01
    cap = rng.uniform(0, 5.2, J); r = rng.uniform(0, 1, J); u = rng.uniform(0, 30, J)
    cap[:8] = 5.0 - r[:8] * u[:8] * dt                   # capacitor lands exactly on the threshold
    free = rng.uniform(0, 12, J); free[8:16] = 5.0
    ln = np.full(J, 5.0)
    w = rng.normal(size=(3, J))
    outs = []
    for d in (None, dev):
        cap_t, r_t, u_t = _leafs([cap, r, u], d)
        co, sp, vn, an = (S.macro_to_micro if d is None else F.macro_to_micro)(
            cap_t, r_t, u_t, torch.tensor(free, dtype=torch.float64, device=d), torch.tensor(ln, dtype=torch.float64, device=d), dt)
        wt = torch.tensor(w, dtype=torch.float64, device=d)
        ((co * wt[0]).sum() + (vn * wt[1]).sum() + (an * wt[2]).sum()).backward()
        outs.append([x.detach().cpu().numpy() for x in (co, sp, vn, an, cap_t.grad, r_t.grad, u_t.grad)])
    a, b = outs
    assert (a[1] == b[1]).all() and 0 < a[1].sum() < J           # spawn decisions identical, both branches taken
    for x, y in zip(a, b):
        assert relerr(y, x) < 1e-14


@pytest.mark.parametrize("N,dx", [(10, 5.0), (12, 2.0), (3, 7.5), (1, 5.0)])
def test_micro_to_macro_batched(dev, N, dx):
    from dhts_b200 import functional as F
    rng = np.random.default_rng(N)
    J, L, umax = 257, 50.0, 30.0
    p = L + rng.uniform(3.0, 9.0, J)                     # some heads past L + len (absorbed), some not
    p[:6] = L + 5.0                                       # exactly at the threshold: not absorbed (strict >)
    v = rng.uniform(0, 30, J); a = rng.uniform(4.0, 5.5, J)
    r = rng.uniform(0, 1, (J, N)); r[::5] = rng.uniform(0.97, 1.2, (len(r[::5]), N))       # upper clamp
    a[1::7] = -rng.uniform(4, 6, len(a[1::7])); r[1::7] *= 0.01                              # lower clamp
    uu = rng.uniform(0, 30, (J, N)); y = r * (uu - umax * (1 - np.sqrt(np.maximum(r, 0) + 1e-5)))
    w = rng.normal(size=(3, J, N))
    const = lambda val, d: torch.full((J,), val, dtype=torch.float64, device=d)
    outs = []
    for d in (None, dev):
        p_t, v_t, a_t, r_t, y_t, u_t = _leafs([p, v, a, r, y, uu], d)
        ro, yo, uo, ab, nt = (S.micro_to_macro if d is None else F.micro_to_macro)(
            p_t, v_t, a_t, const(5.0, d), const(L, d), r_t, y_t, u_t, const(dx, d), const(umax, d))
        wt = torch.tensor(w, dtype=torch.float64, device=d)
        ((ro * wt[0]).sum() + (yo * wt[1]).sum() + (uo * wt[2]).sum()).backward()
        outs.append([x.detach().cpu().numpy() for x in
                     (ro, yo, uo, ab, nt, p_t.grad, v_t.grad, a_t.grad, r_t.grad, y_t.grad, u_t.grad)])
    a_, b_ = outs
    assert (a_[3] == b_[3]).all() and (a_[4] == b_[4]).all()
    assert 0 < a_[3].sum() < J and a_[3][:6].sum() == 0 and a_[4].max() >= 1
    for k, (x, y_) in enumerate(zip(a_, b_)):
        assert relerr(y_, x) < 1e-12, k
