"""TEST INFRASTRUCTURE: freeze outputs of the LIVE reference into tests/golden/.

Run in THIS container only (needs /root/reference, which does not travel to
the GPU box):

    python oracle/gen_golden.py --tier fp64     # dtype-proxied fp64 reference (SURVEY 8c tier 2)
    python oracle/gen_golden.py --tier fp32     # reference as shipped        (SURVEY 8c tier 1)

Each tier must run in its own process (the fp64 switch rebinds module
globals).  The reference is imported unmodified from /root/reference; the
fp64 switch is the recipe of SURVEY.md Appendix C (no file edits).

Fixtures written (small .npz, committed):
  arz_step_<tier>.npz      single-step cases incl. branch-threshold edge cases
  arz_rollout_<tier>.npz   T-step rollout + gradients (states, ghosts)
  idm_step_<tier>.npz      single-step cases incl. clips / tight gaps
  idm_rollout_<tier>.npz   T-step rollout + gradients (states, head deltas)
  hybrid_chain_<tier>.npz  macro->micro->macro chain through RoadNetwork.forward
  inverse_<tier>.npz       seeded Adam loss curves of the three inverse drivers' loop
"""
import argparse
import os
import sys
import time

import numpy as np
import torch as th

REF = os.environ.get("DHTS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def switch_fp64():
    class _NP64:
        float32 = np.float64

        def __getattr__(self, k):
            return getattr(np, k)

    class _TH64:
        float32 = th.float64

        def __getattr__(self, k):
            return getattr(th, k)

    th.set_default_dtype(th.float64)
    import model.macro.darz as a, road.lane.dmacro_lane as b, road.lane._macro_lane as c
    import road.lane.dmicro_lane as d, road.lane._micro_lane as e, model.micro.didm as f
    for m in (a, b, c, d):
        m.np = _NP64()
    for m in (b, c, d, e, f):
        m.th = _TH64()


def f(x):
    return float(x.item()) if isinstance(x, th.Tensor) else float(x)


class Ctx:
    pass


def cell_records(lane):
    cells = [lane.leftmost_cell] + list(lane.curr_cell) + [lane.rightmost_cell]
    return (np.array([f(c.state.q.r) for c in cells]), np.array([f(c.state.q.y) for c in cells]),
            np.array([f(c.state.u) for c in cells]), np.array([f(c.state.u_eq) for c in cells]))


def gen_arz_step(dtype, rng):
    from road.lane.dmacro_lane import dMacroLane, dMacroForwardLayer
    N, dx, umax, dt = 24, 5.0, 30.0, 0.01
    cases = []
    for k in range(24):
        r = rng.uniform(0.0, 1.0, N + 2)
        u = rng.uniform(0.0, 1.0, N + 2) * umax
        if k % 4 == 1:      # near-vacuum cells on either side of interfaces
            idx = rng.choice(N + 2, 6, replace=False)
            r[idx] = rng.choice([0.0, 5e-6, 2e-5], 6)
        if k % 4 == 2:      # equal speeds / small speed differences
            idx = rng.choice(N + 1, 6, replace=False)
            u[idx + 1] = u[idx] + rng.choice([0.0, 5e-6, -5e-6, 2e-5], 6)
        if k % 4 == 3:      # strong rarefactions / vacuum-forming (u_R >> u_L) and congested shocks
            idx = rng.choice(N + 1, 8, replace=False)
            u[idx] = rng.uniform(0, 2, 8); u[idx + 1] = rng.uniform(25, 45, 8)
            r[idx] = rng.uniform(0.001, 0.3, 8)
        if k >= 16:         # congested: negative characteristic speeds -> Q_M / Q_C outcomes
            r = rng.uniform(0.5, 1.2, N + 2)
            u = rng.uniform(0.0, 0.3, N + 2) * umax
        lane = dMacroLane(0, N * dx, umax, dx)
        tr = th.tensor(r, dtype=dtype); tu = th.tensor(u, dtype=dtype)
        lane.set_state_vector_u(tr[1:-1], tu[1:-1])
        lane.set_leftmost_cell(tr[0], tu[0])
        lane.set_rightmost_cell(tr[-1], tu[-1])
        pr, py, pu, pe = cell_records(lane)
        lane.forward(dt)
        nr, ny, nu = lane.get_next_state_vector()
        case = np.array([s.case_ind for s in lane.riemann_solution], dtype=np.int32)
        speeds = np.array([[f(s.speed0), f(s.speed1)] for s in lane.riemann_solution])
        dqs = np.array(lane.d_lane[-1].dqs, dtype=np.float64)
        g_nr = rng.normal(size=N); g_ny = rng.normal(size=N)
        ctx = Ctx(); ctx.dqs = lane.d_lane[-1].dqs
        _, g_r, g_y, _ = dMacroForwardLayer.backward(ctx, th.tensor(g_nr, dtype=dtype), th.tensor(g_ny, dtype=dtype))
        cases.append(dict(pr=pr, py=py, pu=pu, pe=pe, nr=nr.detach().numpy().astype(np.float64),
                          ny=ny.detach().numpy().astype(np.float64), nu=nu.detach().numpy().astype(np.float64),
                          case=case, speeds=speeds, dqs=dqs, g_nr=g_nr, g_ny=g_ny,
                          g_r=g_r.numpy().astype(np.float64), g_y=g_y.numpy().astype(np.float64)))
    out = {k: np.stack([c[k] for c in cases]) for k in cases[0]}
    out.update(N=N, dx=dx, umax=umax, dt=dt)
    hist = np.bincount(out["case"].ravel(), minlength=3)
    print("arz_step: case histogram", hist)
    return out


def gen_arz_rollout(dtype, rng, T):
    from road.lane.dmacro_lane import dMacroLane
    from road.network.road_network import RoadNetwork
    B, N, dx, umax, dt = 3, 32, 5.0, 30.0, 0.01
    r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 1, (B, N)) * umax
    gh = np.stack([rng.uniform(0, 1, (B, 2)), rng.uniform(0, 1, (B, 2)) * umax], axis=-1)  # [B,2,(r,u)]
    w_r = rng.normal(size=(B, N)); w_u = rng.normal(size=(B, N)) / umax
    snaps = sorted(set([1, 2, 10, T // 2, T]))
    res = dict(rT=[], yT=[], uT=[], g_r0=[], g_u0=[], g_ghost=[], snap=[])
    for b in range(B):
        tr = th.tensor(r0[b], dtype=dtype, requires_grad=True)
        tu = th.tensor(u0[b], dtype=dtype, requires_grad=True)
        tg = th.tensor(gh[b], dtype=dtype, requires_grad=True)
        lane = dMacroLane(0, N * dx, umax, dx)
        lane.set_state_vector_u(tr, tu)
        lane.set_leftmost_cell(tg[0, 0], tg[0, 1])
        lane.set_rightmost_cell(tg[1, 0], tg[1, 1])
        net = RoadNetwork(umax)
        net.add_lane(lane)
        sn = []
        for t in range(T):
            net.forward(dt, True)
            if t + 1 in snaps:
                s = lane.get_state_vector()
                sn.append(np.stack([s[0].detach().numpy(), s[1].detach().numpy()], -1).astype(np.float64))
        r, y, u = lane.get_state_vector()
        loss = (r * th.tensor(w_r[b], dtype=dtype)).sum() + (u * th.tensor(w_u[b], dtype=dtype)).sum()
        loss.backward()
        res["rT"].append(r.detach().numpy().astype(np.float64)); res["yT"].append(y.detach().numpy().astype(np.float64))
        res["uT"].append(u.detach().numpy().astype(np.float64))
        res["g_r0"].append(tr.grad.numpy().astype(np.float64)); res["g_u0"].append(tu.grad.numpy().astype(np.float64))
        res["g_ghost"].append(tg.grad.numpy().astype(np.float64)); res["snap"].append(np.stack(sn))
    out = {k: np.stack(v) for k, v in res.items()}
    out.update(r0=r0, u0=u0, ghost_ru=gh, w_r=w_r, w_u=w_u, snaps=np.array(snaps), B=B, N=N, dx=dx, umax=umax, dt=dt,
               T=T)
    return out


def rand_params(rng, n, umax, length=5.0, uniform=False):
    if uniform:   # road/vehicle/micro_vehicle.py:30-72
        return np.stack([np.full(n, umax * 1.0), np.full(n, umax * 0.8), np.full(n, umax * 0.9),
                         np.full(n, length * 0.1), np.full(n, 0.1), np.full(n, length)])
    # road/vehicle/micro_vehicle.py:74-122 ranges
    return np.stack([rng.uniform(1.5, 2.0, n) * umax, rng.uniform(1.0, 1.5, n) * umax, rng.uniform(0.8, 1.2, n) * umax,
                     rng.uniform(0.2, 0.4, n) * length, rng.uniform(0.2, 0.6, n), np.full(n, length)])


def make_micro_lane(p, v, params, dtype, head_dp, head_dv, requires_grad=False):
    from road.lane.dmicro_lane import dMicroLane
    from road.vehicle.micro_vehicle import MicroVehicle
    n = len(p)
    lane = dMicroLane(0, 1e10, 30.0)
    tp = th.tensor(p, dtype=dtype, requires_grad=requires_grad)
    tv = th.tensor(v, dtype=dtype, requires_grad=requires_grad)
    for i in range(n):
        mv = MicroVehicle(i, tp[i], tv[i], float(params[0, i]), float(params[1, i]), float(params[2, i]),
                          float(params[3, i]), float(params[4, i]), float(params[5, i]), float(params[5, i]))
        lane.add_head_vehicle(mv)
    lane.head_position_delta = head_dp
    lane.head_speed_delta = head_dv
    return lane, tp, tv


def gen_idm_step(dtype, rng):
    from road.lane.dmicro_lane import dMicroForwardLayer
    n, umax, dt = 20, 30.0, 0.01
    cases = []
    for k in range(16):
        params = rand_params(rng, n, umax, uniform=(k % 4 == 3))
        if k % 2 == 0:   # throughput-config spacing
            p = np.arange(n) * 20.0 + rng.uniform(0, 1, n) * 10
            v = rng.uniform(9, 21, n)
        else:            # tight gaps: acceleration clip / s* clip exercised
            p = np.cumsum(rng.uniform(5.05, 9.0, n))
            v = rng.uniform(0, 25, n)
            v[rng.choice(n, 4, replace=False)] = rng.uniform(0, 0.05, 4)
        if k % 4 == 1:   # leader much faster: s* < 0 clip
            v[::2] = rng.uniform(0.5, 3, len(v[::2])); v[1::2] = rng.uniform(40, 60, len(v[1::2]))
            params[3] = 0.01; params[4] = 0.001
        head_dp, head_dv = (1000.0, 0.0) if k % 3 else (float(rng.uniform(5, 40)), float(rng.uniform(-5, 5)))
        lane, _, _ = make_micro_lane(p, v, params, dtype, head_dp, head_dv)
        lane.forward(dt)
        np_, nv_ = lane.get_next_state_vector()
        flags = np.array([int(a[2]) | (int(a[3]) << 1) for a in lane.acc_info], dtype=np.int32)
        dqs = np.array(lane.d_lane[-1].dqs, dtype=np.float64)
        g_np = rng.normal(size=n); g_ns = rng.normal(size=n)
        ctx = Ctx(); ctx.dqs = lane.d_lane[-1].dqs
        _, g_p, g_s, _ = dMicroForwardLayer.backward(ctx, th.tensor(g_np, dtype=dtype), th.tensor(g_ns, dtype=dtype))
        # inputs as the lane saw them (tensor dtype rounding included)
        p_in = np.array([f(x.position) for x in lane.curr_vehicle]); v_in = np.array([f(x.speed) for x in lane.curr_vehicle])
        cases.append(dict(p=p_in, v=v_in, params=params, head=np.array([head_dp, head_dv]),
                          np=np_.detach().numpy().astype(np.float64), nv=nv_.detach().numpy().astype(np.float64),
                          flags=flags, dqs=dqs, g_np=g_np, g_ns=g_ns, g_p=g_p.numpy().astype(np.float64),
                          g_s=g_s.numpy().astype(np.float64)))
    out = {k: np.stack([c[k] for c in cases]) for k in cases[0]}
    out.update(n=n, dt=dt)
    print("idm_step: flag histogram", np.bincount(out["flags"].ravel(), minlength=4))
    return out


def gen_idm_rollout(dtype, rng, T):
    n, umax, dt = 16, 30.0, 0.01
    L = 4
    res = dict(p0=[], v0=[], params=[], head=[], pT=[], vT=[], g_p0=[], g_v0=[], g_head=[], w_p=[], w_v=[], snap=[])
    snaps = sorted(set([1, 2, 10, T // 2, T]))
    for l in range(L):
        uniform = (l == 3)
        params = rand_params(rng, n, umax, uniform=uniform)
        if l % 2 == 0:
            p = np.arange(n) * 20.0 + rng.uniform(0, 1, n) * 10; v = rng.uniform(9, 21, n)
        else:
            p = np.cumsum(rng.uniform(8.0, 14.0, n)); v = rng.uniform(0, 25, n)
        head = (1000.0, 0.0) if l < 2 else (float(rng.uniform(20, 60)), float(rng.uniform(-2, 2)))
        hd = th.tensor(head, dtype=dtype, requires_grad=True)
        lane, tp, tv = make_micro_lane(p, v, params, dtype, hd[0], hd[1], requires_grad=True)
        w_p = rng.normal(size=n); w_v = rng.normal(size=n)
        sn = []
        for t in range(T):
            lane.forward(dt); lane.update_state()
            if t + 1 in snaps:
                s = lane.get_state_vector()
                sn.append(np.stack([s[0].detach().numpy(), s[1].detach().numpy()], -1).astype(np.float64))
        pT, vT = lane.get_state_vector()
        loss = (pT * th.tensor(w_p, dtype=dtype)).sum() + (vT * th.tensor(w_v, dtype=dtype)).sum()
        loss.backward()
        res["p0"].append(tp.detach().numpy().astype(np.float64)); res["v0"].append(tv.detach().numpy().astype(np.float64))
        res["params"].append(params); res["head"].append(np.array(head))
        res["pT"].append(pT.detach().numpy().astype(np.float64)); res["vT"].append(vT.detach().numpy().astype(np.float64))
        res["g_p0"].append(tp.grad.numpy().astype(np.float64)); res["g_v0"].append(tv.grad.numpy().astype(np.float64))
        res["g_head"].append(hd.grad.numpy().astype(np.float64)); res["w_p"].append(w_p); res["w_v"].append(w_v)
        res["snap"].append(np.stack(sn))
    out = {k: np.stack(v) for k, v in res.items()}
    out.update(n=n, L=L, dt=dt, T=T, snaps=np.array(snaps))
    return out


def gen_hybrid_chain(dtype, rng, T):
    """macro(10) -> micro -> macro(10) through RoadNetwork.forward + Conversion.*
    (SURVEY A.5 probe): loss on lane 2 only, gradient wrt lane 0 initial state."""
    from road.lane.dmacro_lane import dMacroLane
    from road.lane.dmicro_lane import dMicroLane
    from road.network.road_network import RoadNetwork
    N, dx, umax, dt = 10, 5.0, 30.0, 0.01
    Llen = N * dx
    r0 = rng.uniform(0.3, 1.0, (2, N)); u0 = rng.uniform(0.3, 1.0, (2, N)) * umax
    gh = np.stack([rng.uniform(0.3, 1, 4), rng.uniform(0.3, 1, 4) * umax], -1)   # [4,(r,u)]
    np.random.seed(1234)
    net = RoadNetwork(umax)
    t0r = th.tensor(r0[0], dtype=dtype, requires_grad=True); t0u = th.tensor(u0[0], dtype=dtype, requires_grad=True)
    t2r = th.tensor(r0[1], dtype=dtype, requires_grad=True); t2u = th.tensor(u0[1], dtype=dtype, requires_grad=True)
    l0 = dMacroLane(0, Llen, umax, dx); l0.set_state_vector_u(t0r, t0u)
    l0.set_leftmost_cell(th.tensor(gh[0, 0], dtype=dtype), th.tensor(gh[0, 1], dtype=dtype))
    l0.set_rightmost_cell(th.tensor(gh[1, 0], dtype=dtype), th.tensor(gh[1, 1], dtype=dtype))
    net.add_lane(l0)
    l1 = dMicroLane(1, Llen, umax); net.add_lane(l1)
    l2 = dMacroLane(2, Llen, umax, dx); l2.set_state_vector_u(t2r, t2u)
    l2.set_leftmost_cell(th.tensor(gh[2, 0], dtype=dtype), th.tensor(gh[2, 1], dtype=dtype))
    l2.set_rightmost_cell(th.tensor(gh[3, 0], dtype=dtype), th.tensor(gh[3, 1], dtype=dtype))
    net.add_lane(l2)
    net.connect_lane(0, 1); net.connect_lane(1, 2)
    net.macro_route = net.create_random_macro_route()
    nveh_hist = []; nspawn = []; cap_hist = []
    for t in range(T):
        net.forward(dt, True)
        nveh_hist.append(l1.num_vehicle()); nspawn.append(net.num_vehicle)
        cap_hist.append(f(l0.flux_capacitor.get(1, 0.0)))
    s0 = l0.get_state_vector(); s2 = l2.get_state_vector()
    w_r = rng.normal(size=N); w_u = rng.normal(size=N) / umax
    loss = (s2[0] * th.tensor(w_r, dtype=dtype)).sum() + (s2[2] * th.tensor(w_u, dtype=dtype)).sum()
    # vehicles still on the micro lane contribute too (exercise the IDM part of the chain)
    w_veh = rng.normal(size=8)
    for i, mv in enumerate(l1.curr_vehicle):
        loss = loss + w_veh[2 * i % 8] * mv.position * 0.01 + w_veh[(2 * i + 1) % 8] * mv.speed * 0.01
    loss.backward()
    z = lambda t: np.zeros(N) if t.grad is None else t.grad.numpy().astype(np.float64)
    out = dict(r0=r0, u0=u0, ghost_ru=gh, w_r=w_r, w_u=w_u, w_veh=w_veh, N=N, dx=dx, umax=umax, dt=dt, T=T,
               nveh_hist=np.array(nveh_hist), nspawn_hist=np.array(nspawn), cap_hist=np.array(cap_hist),
               lane0=np.stack([x.detach().numpy().astype(np.float64) for x in s0]),
               lane2=np.stack([x.detach().numpy().astype(np.float64) for x in s2]),
               veh=np.array([[f(mv.position), f(mv.speed), f(mv.a)] for mv in l1.curr_vehicle]).reshape(-1, 3),
               loss=f(loss), g_r0_lane0=z(t0r), g_u0_lane0=z(t0u), g_r0_lane2=z(t2r), g_u0_lane2=z(t2u))
    print("hybrid_chain: spawned", net.num_vehicle, "on-lane", l1.num_vehicle(), "|g lane0|",
          np.abs(out["g_r0_lane0"]).max(), np.abs(out["g_u0_lane0"]).max())
    return out


def gen_inverse(dtype, rng, T, episodes):
    """Seeded restatement of the solve_gd loop (example/inverse/_inverse.py:185-242)
    over the three inverse problems' networks (example/inverse/{macro,micro,hybrid}.py),
    using the reference's lanes/network.  Records per-episode end errors."""
    from road.lane.dmacro_lane import dMacroLane
    from road.lane.dmicro_lane import dMicroLane
    from road.network.road_network import RoadNetwork
    from road.network.route import MicroRoute
    from road.vehicle.micro_vehicle import MicroVehicle
    N, dx, umax, dt = 10, 5.0, 30.0, 0.01
    out = {}

    def clear(net):   # _inverse.py:358-370
        for lane in net.lane.values():
            lane.clear()
        net.vehicle.clear(); net.micro_route.clear(); net.num_vehicle = 0

    # ---- macro (macro.py:34-68, 226-241) and hybrid (hybrid.py:37-82): loss on lane 0's (r,u)
    for mode in ("macro", "hybrid"):
        np.random.seed(7)
        nb = 2 if mode == "macro" else 4
        bd = rng.uniform(0, 1, nb); bs = rng.uniform(0, 1, nb) * umax
        true_r = rng.uniform(0, 1, N); true_u = rng.uniform(0, 1, N) * umax
        est_r = rng.uniform(0, 1, N); est_u = rng.uniform(0, 1, N) * umax
        net = RoadNetwork(umax)
        lane = dMacroLane(0, N * dx, umax, dx)
        lane.set_leftmost_cell(th.tensor(bd[0], dtype=dtype), th.tensor(bs[0], dtype=dtype))
        lane.set_rightmost_cell(th.tensor(bd[1], dtype=dtype), th.tensor(bs[1], dtype=dtype))
        net.add_lane(lane)
        if mode == "hybrid":
            net.add_lane(dMicroLane(1, N * dx, umax))
            l2 = dMacroLane(2, N * dx, umax, dx)
            l2.set_leftmost_cell(th.tensor(bd[2], dtype=dtype), th.tensor(bs[2], dtype=dtype))
            l2.set_rightmost_cell(th.tensor(bd[3], dtype=dtype), th.tensor(bs[3], dtype=dtype))
            net.add_lane(l2)
            net.connect_lane(0, 1); net.connect_lane(1, 2)
            net.macro_route = net.create_random_macro_route()
        lane.set_state_vector_u(th.tensor(true_r, dtype=dtype), th.tensor(true_u, dtype=dtype))
        for _ in range(T):
            net.forward(dt, False)
        s = net.lane[0].get_state_vector()
        end_r, end_u = s[0].detach().clone(), s[2].detach().clone()
        er = th.tensor(est_r, dtype=dtype, requires_grad=True); eu = th.tensor(est_u, dtype=dtype, requires_grad=True)
        opt = th.optim.Adam((er, eu), lr=1e-3)
        errs = []
        for _ in range(episodes):
            clear(net)
            net.lane[0].set_state_vector_u(er, eu)
            for _ in range(T):
                net.forward(dt, True)
            s = net.lane[0].get_state_vector()
            err = th.pow(end_r - s[0], 2.0).sum() + th.pow(end_u - s[2], 2.0).sum()
            errs.append(err.item())
            opt.zero_grad(); err.backward(); opt.step()
            with th.no_grad():
                er.clamp_(0.0, 1.0); eu.clamp_(0.0, umax)
        out.update({mode + "_bd": bd, mode + "_bs": bs, mode + "_true_r": true_r, mode + "_true_u": true_u,
                    mode + "_est_r": est_r, mode + "_est_u": est_u, mode + "_end_r": end_r.numpy().astype(np.float64),
                    mode + "_end_u": end_u.numpy().astype(np.float64), mode + "_errs": np.array(errs),
                    mode + "_final_r": er.detach().numpy().astype(np.float64),
                    mode + "_final_u": eu.detach().numpy().astype(np.float64)})
        print("inverse", mode, errs)

    # ---- micro (micro.py:36-118, 221-236), gd_lr = 1e-2 (micro.py:264)
    n, vl = 10, 5.0

    def set_state(net, p, v):
        net.vehicle.clear(); net.micro_route.clear()
        lane = net.lane[0]; lane.clear()
        for i in range(n):
            mv = MicroVehicle.default_micro_vehicle(umax)
            mv.position = p[i]; mv.speed = v[i]
            net.add_vehicle(mv, MicroRoute([0]))
        lane.set_state_vector(p, v)

    true_p = np.arange(n) * 4.0 * vl + rng.uniform(0, 1, n) * 2.0 * vl
    true_v = (0.3 + 0.4 * rng.uniform(0, 1, n)) * umax
    est_p = true_p + rng.normal(size=n) * 0.1 * vl
    est_v = true_v + rng.normal(size=n) * 1e-2 * umax
    net = RoadNetwork(umax)
    net.add_lane(dMicroLane(0, 1e10, umax))
    set_state(net, th.tensor(true_p, dtype=dtype), th.tensor(true_v, dtype=dtype))
    for _ in range(T):
        net.forward(dt, False)
    end_p, end_v = [x.detach().clone() for x in net.lane[0].get_state_vector()]
    ep = th.tensor(est_p, dtype=dtype, requires_grad=True); ev = th.tensor(est_v, dtype=dtype, requires_grad=True)
    opt = th.optim.Adam((ep, ev), lr=1e-2)
    errs = []
    for _ in range(episodes):
        clear(net)
        set_state(net, ep, ev)
        for _ in range(T):
            net.forward(dt, True)
        s = net.lane[0].get_state_vector()
        err = th.pow(end_p - s[0], 2.0).sum() + th.pow(end_v - s[1], 2.0).sum()
        errs.append(err.item())
        opt.zero_grad(); err.backward(); opt.step()
    out.update(micro_true_p=true_p, micro_true_v=true_v, micro_est_p=est_p, micro_est_v=est_v,
               micro_end_p=end_p.numpy().astype(np.float64), micro_end_v=end_v.numpy().astype(np.float64),
               micro_errs=np.array(errs), micro_final_p=ep.detach().numpy().astype(np.float64),
               micro_final_v=ev.detach().numpy().astype(np.float64))
    print("inverse micro", errs)
    out.update(N=N, n=n, dx=dx, umax=umax, dt=dt, T=T, episodes=episodes)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tier", choices=["fp32", "fp64"], required=True)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    if args.tier == "fp64":
        switch_fp64()
        dtype = th.float64
    else:
        dtype = th.float32
    os.makedirs(OUT, exist_ok=True)
    jobs = dict(
        arz_step=lambda: gen_arz_step(dtype, np.random.default_rng(101)),
        arz_rollout=lambda: gen_arz_rollout(dtype, np.random.default_rng(102), 300),
        idm_step=lambda: gen_idm_step(dtype, np.random.default_rng(103)),
        idm_rollout=lambda: gen_idm_rollout(dtype, np.random.default_rng(104), 300),
        hybrid_chain=lambda: gen_hybrid_chain(dtype, np.random.default_rng(105), 700),
        inverse=lambda: gen_inverse(dtype, np.random.default_rng(106), 200, 3),
    )
    for name, fn in jobs.items():
        if args.only and name not in args.only.split(","):
            continue
        t0 = time.time()
        th.manual_seed(20221008); np.random.seed(20221008)
        data = fn()
        np.savez_compressed(os.path.join(OUT, "%s_%s.npz" % (name, args.tier)), **data)
        print("wrote %s_%s.npz in %.1fs" % (name, args.tier, time.time() - t0))


if __name__ == "__main__":
    main()
