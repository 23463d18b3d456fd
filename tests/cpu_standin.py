"""TEST INFRASTRUCTURE ONLY: lets the HOST logic of the drop-in object API (cell / vehicle marshalling, boundary
rules, event handling of the conversions, autograd wiring) run on a box without a GPU by swapping the four
kernel-backed functions of ``dhts_b200.functional`` for CPU stand-ins built on the oracle (oracle/dhts_oracle.c)
and plain torch.  Nothing here is reachable from the product: the patch is applied by a pytest fixture, the
product path itself still refuses to run without CUDA (tests/test_cabi.py, tests/test_dropin_host.py).
"""
import contextlib

import numpy as np
import torch

from oracle import oracle as O

EPS = 1e-5


class _ArzStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r_pad, y_pad, u_pad, ueq_pad, dx, umax, dt, flags_t):
        B = r_pad.shape[0]
        nr, ny, cases, dqs = [], [], [], []
        for b in range(B):
            o = O.arz_step(r_pad[b].numpy(), y_pad[b].numpy(), u_pad[b].numpy(), ueq_pad[b].numpy(), float(dx[b]),
                           float(umax[b]), dt)
            nr.append(o["nr"]); ny.append(o["ny"]); cases.append(o["case"]); dqs.append(o["dqs"])
            if o["cfl"]:
                flags_t[0] |= 1
        ctx.dqs = dqs
        ctx.dtype = r_pad.dtype
        case = torch.tensor(np.stack(cases), dtype=torch.int32)
        ctx.mark_non_differentiable(case)
        return torch.tensor(np.stack(nr), dtype=r_pad.dtype), torch.tensor(np.stack(ny), dtype=r_pad.dtype), case

    @staticmethod
    def backward(ctx, g_nr, g_ny, _):
        gr, gy = [], []
        for b, dq in enumerate(ctx.dqs):
            a, c = O.arz_vjp(dq, g_nr[b].numpy(), g_ny[b].numpy())
            gr.append(a); gy.append(c)
        return (torch.tensor(np.stack(gr), dtype=ctx.dtype), torch.tensor(np.stack(gy), dtype=ctx.dtype), None, None,
                None, None, None, None)


def arz_step(r_pad, y_pad, u_pad, dx, umax, dt, flags, ueq_pad=None, want_case=False):
    B = r_pad.shape[0]
    per = lambda x: torch.as_tensor(x, dtype=torch.float64).expand(B) if torch.as_tensor(x).dim() == 0 else torch.as_tensor(x)
    dxl, uml = per(dx), per(umax)
    if ueq_pad is None:
        ueq_pad = uml[:, None] * (1.0 - torch.sqrt(torch.clamp(r_pad.detach(), min=0.0) + EPS))
    nr, ny, case = _ArzStep.apply(r_pad, y_pad, u_pad.detach(), ueq_pad.detach(), dxl, uml, float(dt), flags.t)
    rc = torch.clamp(nr, min=EPS)                     # compute_u with its true derivative (set_r_y, _arz.py:88-92)
    nu = ny / rc + uml[:, None].to(nr.dtype) * (1.0 - torch.sqrt(rc + EPS))
    return (nr, ny, nu, case) if want_case else (nr, ny, nu)


class _IdmStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, v, head, params, lane_off, dt, flags_t):
        off = lane_off.tolist()
        np_, nv_, vf = torch.empty_like(p), torch.empty_like(p), torch.zeros(p.numel(), dtype=torch.int32)
        ctx.dqs, ctx.off, ctx.dtype = [], off, p.dtype
        for l in range(len(off) - 1):
            a, b = off[l], off[l + 1]
            if b == a:
                ctx.dqs.append(None)
                continue
            o = O.idm_step(p[a:b].numpy(), v[a:b].numpy(), params[:, a:b].numpy(), float(head[l, 0]), float(head[l, 1]), dt)
            np_[a:b] = torch.tensor(o["np"], dtype=p.dtype); nv_[a:b] = torch.tensor(o["nv"], dtype=p.dtype)
            vf[a:b] = torch.tensor(o["flags"], dtype=torch.int32)
            ctx.dqs.append(o["dqs"])
            if o["ncol"]:
                flags_t[0] |= 4
                flags_t[1] += int(o["ncol"])
        ctx.mark_non_differentiable(vf)
        return np_, nv_, vf

    @staticmethod
    def backward(ctx, g_np, g_nv, _):
        off = ctx.off
        g_p, g_v = torch.zeros_like(g_np), torch.zeros_like(g_np)
        g_head = torch.zeros((len(off) - 1, 2), dtype=ctx.dtype)
        for l, dq in enumerate(ctx.dqs):
            a, b = off[l], off[l + 1]
            if dq is None:
                continue
            gp, gs = O.idm_vjp(dq, g_np[a:b].numpy(), g_nv[a:b].numpy())
            gp, gs = torch.tensor(gp, dtype=ctx.dtype), torch.tensor(gs, dtype=ctx.dtype)
            g_p[a:b], g_v[a:b] = gp[:-1], gs[:-1]
            g_p[b - 1] += gp[-1]; g_v[b - 1] += gs[-1]       # ghost = (p_head + dp, v_head - dv), dmicro_lane.py:144-151
            g_head[l, 0], g_head[l, 1] = gp[-1], -gs[-1]
        return g_p, g_v, g_head, None, None, None, None


def idm_step(p, v, params, lane_off, head, dt, flags, veh_lane=None, want_flags=False):
    np_, nv_, vf = _IdmStep.apply(p, v, head, params.detach(), lane_off, float(dt), flags.t)
    return (np_, nv_, vf) if want_flags else (np_, nv_)


def macro_to_micro(cap, r_last, u_last, free_space, veh_len, dt):
    """conversion.py:32-68 in torch (autograd gives the adjoint)."""
    flux = cap + r_last * u_last * dt
    sp = (flux >= veh_len) & (free_space >= veh_len)
    zero = torch.zeros_like(flux)
    a_new = torch.where(sp, flux - (flux.detach() - veh_len), zero)
    v_new = torch.where(sp, u_last, zero)
    cap_out = torch.where(sp, (flux - veh_len).detach(), flux)
    return cap_out, sp.to(torch.int32), v_new, a_new


def micro_to_macro(p_head, v_head, a_head, len_head, lane_len, r, y, u, dx, umax):
    """conversion.py:99-171 in torch, one junction at a time."""
    J, N = r.shape
    r_out, y_out, u_out = [], [], []
    absorbed = torch.zeros(J, dtype=torch.int32); ntouched = torch.zeros(J, dtype=torch.int32)
    for j in range(J):
        rr, yy, uu = list(r[j].unbind(0)), list(y[j].unbind(0)), list(u[j].unbind(0))
        ln, d, um = len_head[j], dx[j], umax[j]
        if float(p_head[j]) > float(lane_len[j]) + float(ln):
            absorbed[j] = 1
            vh = p_head[j] - lane_len[j]; vt = vh - ln
            for ci in range(N):
                c_head, c_tail = d * (ci + 1), d * ci
                if not (float(c_head) > float(vt) and float(c_tail) < float(vh)):
                    break
                max_head = c_head if float(c_head) > float(vh) else vh
                min_tail = c_tail if float(c_tail) < float(vt) else vt
                overlap = d + ln - (max_head - min_tail)
                n_r = rr[ci] + (a_head[j] / ln.detach()) * (overlap / d)
                val = float(n_r)
                if val > 1.0 - 1e-5:
                    n_r = n_r - (val - (1.0 - 1e-5))
                elif val < 1e-5:
                    n_r = n_r - (val - 1e-5)
                rr[ci] = n_r; uu[ci] = v_head[j]
                yy[ci] = n_r * (v_head[j] - um * (1.0 - torch.sqrt(torch.clamp(n_r, min=0.0) + EPS)))
                ntouched[j] += 1
        r_out.append(torch.stack(rr)); y_out.append(torch.stack(yy)); u_out.append(torch.stack(uu))
    return torch.stack(r_out), torch.stack(y_out), torch.stack(u_out), absorbed, ntouched


def arz_rollout_state(r0, y0, u0, ghost, dx, umax, dt, steps, ckpt_every, flags):
    """Stand-in of the fused ARZ rollout operator: the one-step stand-in chained `steps` times over static ghosts."""
    r, y, u = r0, y0, u0.detach()
    ue = None
    for _ in range(int(steps)):
        pad = lambda x, k: torch.cat([ghost[:, 0:1, k], x, ghost[:, 1:2, k]], dim=1)
        r, y, u = arz_step(pad(r, 0), pad(y, 1), pad(u, 2).detach(), dx, umax, dt, flags)
    return r, y, u


def idm_rollout_state(p0, v0, params, lane_off, head, dt, steps, ckpt_every, flags, max_lane):
    p, v = p0, v0
    for _ in range(int(steps)):
        p, v = idm_step(p, v, params, lane_off, head, dt, flags)
    return p, v


@contextlib.contextmanager
def patched(precision="float64"):
    """Route the drop-in lanes through the stand-ins above, on the CPU.  Restores everything on exit."""
    import dhts_b200.dropin as dropin
    from dhts_b200 import functional as F
    from dhts_b200.dropin import runtime as rt
    dropin.install(precision=precision)
    saved = {k: getattr(F, k) for k in ("arz_step", "idm_step", "macro_to_micro", "micro_to_macro", "arz_rollout_state",
                                        "idm_rollout_state")}
    saved_dev, saved_flags = rt.device, dict(rt._flags)
    F.arz_step, F.idm_step, F.macro_to_micro, F.micro_to_macro = arz_step, idm_step, macro_to_micro, micro_to_macro
    F.arz_rollout_state, F.idm_rollout_state = arz_rollout_state, idm_rollout_state
    rt.device = lambda: torch.device("cpu")
    rt._flags.clear()
    saved_hyb = rt._cfg["defer_hyb"]
    rt.configure(defer_hyb=False)          # the fused hybrid rollout has no CPU stand-in: connected networks step immediately
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(F, k, v)
        rt.device = saved_dev
        rt.configure(defer_hyb=saved_hyb)
        rt._flags.clear(); rt._flags.update(saved_flags)
