"""Shared bodies of tests/test_dropin_gpu.py and tests/test_dropin_host.py.

Drop-in object API (road.* / model.* / dmath.* of dhts_b200.dropin) on the GPU against outputs of the LIVE
reference frozen by oracle/gen_golden.py: the loops below restate that script's loops (which in turn restate
example/inverse/_inverse.py:185-242 and the SURVEY A.5 chain probe) over OUR lanes and network.

Tolerances.  precision="float64" vs the dtype-proxied fp64 reference: rtol 1e-5 is the north-star bar; asserted
far tighter (1e-9 .. 1e-8 of the largest entry) because only re-association separates the two.  precision="mixed"
(fp32 storage, fp64 step = the reference as shipped) vs the fp32 reference: states 2e-5, gradients / Adam
iterates 5e-4 of the largest entry (the reference rounds its Jacobians to fp32, we do not).
"""
import numpy as np
import pytest
import torch as th

from conftest import golden, relerr

TIERS = [("fp64", "float64", th.float64, 1e-9, 1e-8), ("fp32", "mixed", th.float32, 2e-5, 5e-4)]


def _install(precision):
    """The GPU tests install the real drop-in; the host tests have already routed it through tests/cpu_standin.py."""
    import dhts_b200.dropin as dropin
    from dhts_b200.dropin import runtime as rt
    dropin.install()
    rt.configure(precision=precision)


def _clear(net):      # example/inverse/_inverse.py:358-370
    for lane in net.lane.values():
        lane.clear()
    net.vehicle.clear(); net.micro_route.clear(); net.num_vehicle = 0


def case_inverse_macro_and_hybrid_loss_curves(tier, precision, dtype, tol_s, tol_g, mode):
    _install(precision)
    from road.lane.dmacro_lane import dMacroLane
    from road.lane.dmicro_lane import dMicroLane
    from road.network.road_network import RoadNetwork
    g = golden("inverse_" + tier)
    N, dx, umax, dt, T, episodes = int(g["N"]), float(g["dx"]), float(g["umax"]), float(g["dt"]), int(g["T"]), int(g["episodes"])
    t = lambda a: th.tensor(a, dtype=dtype)
    np.random.seed(7)
    bd, bs = g[mode + "_bd"], g[mode + "_bs"]
    net = RoadNetwork(umax)
    lane = dMacroLane(0, N * dx, umax, dx)
    lane.set_leftmost_cell(t(bd[0]), t(bs[0])); lane.set_rightmost_cell(t(bd[1]), t(bs[1]))
    net.add_lane(lane)
    if mode == "hybrid":
        net.add_lane(dMicroLane(1, N * dx, umax))
        l2 = dMacroLane(2, N * dx, umax, dx)
        l2.set_leftmost_cell(t(bd[2]), t(bs[2])); l2.set_rightmost_cell(t(bd[3]), t(bs[3]))
        net.add_lane(l2)
        net.connect_lane(0, 1); net.connect_lane(1, 2)
        net.macro_route = net.create_random_macro_route()
    lane.set_state_vector_u(t(g[mode + "_true_r"]), t(g[mode + "_true_u"]))
    for _ in range(T):
        net.forward(dt, False)
    s = net.lane[0].get_state_vector()
    end_r, end_u = s[0].detach().clone(), s[2].detach().clone()
    assert not end_r.is_cuda and end_r.dtype == dtype          # host tensors of the reference's dtype
    assert relerr(end_r, g[mode + "_end_r"]) < tol_s and relerr(end_u, g[mode + "_end_u"]) < tol_s
    er = t(g[mode + "_est_r"]).requires_grad_(); eu = t(g[mode + "_est_u"]).requires_grad_()
    opt = th.optim.Adam((er, eu), lr=1e-3)
    errs = []
    for _ in range(episodes):
        _clear(net)
        net.lane[0].set_state_vector_u(er, eu)
        for _ in range(T):
            net.forward(dt, True)
        s = net.lane[0].get_state_vector()
        err = th.pow(end_r - s[0], 2.0).sum() + th.pow(end_u - s[2], 2.0).sum()
        errs.append(err.item())
        opt.zero_grad(); err.backward(); opt.step()
        with th.no_grad():
            er.clamp_(0.0, 1.0); eu.clamp_(0.0, umax)
    assert relerr(errs, g[mode + "_errs"]) < max(tol_s * 10, 1e-8)
    # Adam's first steps move every entry by ~lr regardless of gradient scale, so the iterates pin the SIGN and
    # relative size of every gradient entry
    assert relerr(er.detach(), g[mode + "_final_r"]) < tol_g and relerr(eu.detach(), g[mode + "_final_u"]) < tol_g


def case_inverse_micro_loss_curve(tier, precision, dtype, tol_s, tol_g):
    _install(precision)
    from road.lane.dmicro_lane import dMicroLane
    from road.network.road_network import RoadNetwork
    from road.network.route import MicroRoute
    from road.vehicle.micro_vehicle import MicroVehicle
    g = golden("inverse_" + tier)
    n, umax, dt, T, episodes = int(g["n"]), float(g["umax"]), float(g["dt"]), int(g["T"]), int(g["episodes"])
    t = lambda a: th.tensor(a, dtype=dtype)

    def set_state(net, p, v):      # example/inverse/micro.py:94-118
        net.vehicle.clear(); net.micro_route.clear()
        lane = net.lane[0]; lane.clear()
        for i in range(n):
            mv = MicroVehicle.default_micro_vehicle(umax)
            mv.position = p[i]; mv.speed = v[i]
            net.add_vehicle(mv, MicroRoute([0]))
        lane.set_state_vector(p, v)

    net = RoadNetwork(umax)
    net.add_lane(dMicroLane(0, 1e10, umax))
    set_state(net, t(g["micro_true_p"]), t(g["micro_true_v"]))
    for _ in range(T):
        net.forward(dt, False)
    end_p, end_v = [x.detach().clone() for x in net.lane[0].get_state_vector()]
    assert relerr(end_p, g["micro_end_p"]) < tol_s and relerr(end_v, g["micro_end_v"]) < tol_s
    ep = t(g["micro_est_p"]).requires_grad_(); ev = t(g["micro_est_v"]).requires_grad_()
    opt = th.optim.Adam((ep, ev), lr=1e-2)
    errs = []
    for _ in range(episodes):
        _clear(net)
        set_state(net, ep, ev)
        for _ in range(T):
            net.forward(dt, True)
        s = net.lane[0].get_state_vector()
        err = th.pow(end_p - s[0], 2.0).sum() + th.pow(end_v - s[1], 2.0).sum()
        errs.append(err.item())
        opt.zero_grad(); err.backward(); opt.step()
    # the loss is a small difference of positions ~1e2: fp32 storage leaves ~1e-4 relative on it
    assert relerr(errs, g["micro_errs"]) < (1e-8 if tier == "fp64" else 1e-3)
    assert relerr(ep.detach(), g["micro_final_p"]) < tol_g and relerr(ev.detach(), g["micro_final_v"]) < tol_g


def case_hybrid_chain_spawn_absorb_and_gradients(tier, precision, dtype, tol_s, tol_g):
    """macro(10) -> micro -> macro(10): vehicles spawned from lane 0's flux capacitor, driven by IDM, absorbed into
    lane 2; loss on lane 2 (+ vehicles still on lane 1) must reach lane 0's initial state."""
    _install(precision)
    from road.lane.dmacro_lane import dMacroLane
    from road.lane.dmicro_lane import dMicroLane
    from road.network.road_network import RoadNetwork
    g = golden("hybrid_chain_" + tier)
    N, dx, umax, dt, T = int(g["N"]), float(g["dx"]), float(g["umax"]), float(g["dt"]), int(g["T"])
    r0, u0, gh = g["r0"], g["u0"], g["ghost_ru"]
    t = lambda a: th.tensor(a, dtype=dtype)
    np.random.seed(1234)
    net = RoadNetwork(umax)
    t0r, t0u = t(r0[0]).requires_grad_(), t(u0[0]).requires_grad_()
    t2r, t2u = t(r0[1]).requires_grad_(), t(u0[1]).requires_grad_()
    l0 = dMacroLane(0, N * dx, umax, dx); l0.set_state_vector_u(t0r, t0u)
    l0.set_leftmost_cell(t(gh[0, 0]), t(gh[0, 1])); l0.set_rightmost_cell(t(gh[1, 0]), t(gh[1, 1]))
    net.add_lane(l0)
    l1 = dMicroLane(1, N * dx, umax); net.add_lane(l1)
    l2 = dMacroLane(2, N * dx, umax, dx); l2.set_state_vector_u(t2r, t2u)
    l2.set_leftmost_cell(t(gh[2, 0]), t(gh[2, 1])); l2.set_rightmost_cell(t(gh[3, 0]), t(gh[3, 1]))
    net.add_lane(l2)
    net.connect_lane(0, 1); net.connect_lane(1, 2)
    net.macro_route = net.create_random_macro_route()
    nveh, nspawn, cap = [], [], []
    for _ in range(T):
        net.forward(dt, True)
        nveh.append(l1.num_vehicle()); nspawn.append(net.num_vehicle)
        c = l0.flux_capacitor.get(1, 0.0)
        cap.append(float(c.detach()) if th.is_tensor(c) else float(c))
    # discrete events identical, step by step
    assert nspawn == g["nspawn_hist"].tolist() and nveh == g["nveh_hist"].tolist()
    assert np.abs(np.array(cap) - g["cap_hist"]).max() < max(tol_s, 1e-9) * 5.0
    s0, s2 = l0.get_state_vector(), l2.get_state_vector()
    for k in range(3):
        assert relerr(s0[k].detach(), g["lane0"][k]) < tol_s * 5, ("lane0", k)
        assert relerr(s2[k].detach(), g["lane2"][k]) < tol_s * 5, ("lane2", k)
    fl = lambda x: float(x.detach()) if th.is_tensor(x) else float(x)
    veh = np.array([[fl(mv.position), fl(mv.speed), fl(mv.a)] for mv in l1.curr_vehicle]).reshape(-1, 3)
    assert veh.shape == g["veh"].shape and relerr(veh, g["veh"]) < tol_s * 5
    loss = (s2[0] * t(g["w_r"])).sum() + (s2[2] * t(g["w_u"])).sum()
    w_veh = g["w_veh"]
    for i, mv in enumerate(l1.curr_vehicle):
        loss = loss + w_veh[2 * i % 8] * mv.position.cpu() * 0.01 + w_veh[(2 * i + 1) % 8] * mv.speed.cpu() * 0.01
    assert abs(float(loss) - float(g["loss"])) < tol_s * 50
    loss.backward()
    for name, leaf in (("g_r0_lane0", t0r), ("g_u0_lane0", t0u), ("g_r0_lane2", t2r), ("g_u0_lane2", t2u)):
        ref = g[name]
        got = np.zeros_like(ref) if leaf.grad is None else leaf.grad.numpy()
        assert relerr(got, ref) < tol_g * (1 if tier == "fp64" else 4), name
    assert np.abs(g["g_r0_lane0"]).max() > 0        # the spawn/absorb adjoint path is live


def case_macro_lane_rollout_and_ghost_gradients(tier, precision, dtype, tol_s, tol_g):
    """Single lanes through RoadNetwork.forward with differentiable ghost cells (the ITSCP control-gradient path)."""
    _install(precision)
    from road.lane.dmacro_lane import dMacroLane
    from road.network.road_network import RoadNetwork
    g = golden("arz_rollout_" + tier)
    B, N, dx, umax, dt, T = int(g["B"]), int(g["N"]), float(g["dx"]), float(g["umax"]), float(g["dt"]), int(g["T"])
    for b in range(B):
        tr = th.tensor(g["r0"][b], dtype=dtype, requires_grad=True)
        tu = th.tensor(g["u0"][b], dtype=dtype, requires_grad=True)
        tg = th.tensor(g["ghost_ru"][b], dtype=dtype, requires_grad=True)
        lane = dMacroLane(0, N * dx, umax, dx)
        lane.set_state_vector_u(tr, tu)
        lane.set_leftmost_cell(tg[0, 0], tg[0, 1]); lane.set_rightmost_cell(tg[1, 0], tg[1, 1])
        net = RoadNetwork(umax); net.add_lane(lane)
        for _ in range(T):
            net.forward(dt, True)
        r, y, u = lane.get_state_vector()
        assert relerr(r.detach(), g["rT"][b]) < tol_s * 10 and relerr(u.detach(), g["uT"][b]) < tol_s * 10
        loss = (r * th.tensor(g["w_r"][b], dtype=dtype)).sum() + (u * th.tensor(g["w_u"][b], dtype=dtype)).sum()
        loss.backward()
        assert relerr(tr.grad, g["g_r0"][b]) < tol_g and relerr(tu.grad, g["g_u0"][b]) < tol_g
        assert relerr(tg.grad, g["g_ghost"][b]) < tol_g


def case_object_surface_on_device():
    """Cells / vehicles are assignable records; a caller rewriting them is honoured on the next step."""
    _install("float64")
    from road.lane.dmacro_lane import dMacroLane, dMacroForwardLayer
    from road.lane.dmicro_lane import dMicroLane, dMicroForwardLayer
    from road.vehicle.micro_vehicle import MicroVehicle
    import copy
    lane = dMacroLane(0, 50.0, 30.0, 5.0)
    lane.set_state_vector_u(th.rand(10, dtype=th.float64), th.rand(10, dtype=th.float64) * 30)
    lane.set_leftmost_cell(0.3, 12.0); lane.set_rightmost_cell(0.4, 9.0)
    twin = copy.deepcopy(lane)                       # trainer.py:172 deep-copies the environment
    lane.forward(0.01); lane.update_state()
    a = lane.get_state_vector()[0].clone()
    twin.curr_cell[3].state.q.r = 0.9                # plain float written by a caller
    twin.forward(0.01); twin.update_state()
    b = twin.get_state_vector()[0]
    assert (a - b).abs().max() > 1e-4 and (a - b)[:2].abs().max() == 0 and (a - b)[6:].abs().max() == 0
    cr, cy = lane.vectorize_input()
    nr, ny = dMacroForwardLayer.apply(lane, cr, cy, 0.01)
    assert nr.shape == (10,) and cr.shape == (12,)
    ml = dMicroLane(1, 1e4, 30.0)
    for i in range(5):
        mv = MicroVehicle.default_micro_vehicle(30.0); mv.position = 30.0 * i; mv.speed = 10.0 + i
        ml.add_head_vehicle(mv)
    ml.forward(0.01); ml.update_state()
    cp, cs = ml.vectorize_input()
    assert cp.shape == (6,) and abs(float(cp[-1] - cp[-2]) - 1000.0) < 1e-9
    np_, ns_ = dMicroForwardLayer.apply(ml, cp, cs, 0.01)
    ml.forward(0.01)
    p2, s2 = ml.get_next_state_vector()
    assert relerr(np_.cpu(), p2) < 1e-14 and relerr(ns_.cpu(), s2) < 1e-12


def case_cfl_violation_raises_like_the_reference():
    _install("float64")
    from road.lane.dmacro_lane import dMacroLane
    from road.network.road_network import RoadNetwork
    lane = dMacroLane(0, 1.0, 30.0, 0.1)
    lane.set_state_vector_u(th.full((10,), 0.5, dtype=th.float64), th.full((10,), 29.0, dtype=th.float64))
    lane.set_leftmost_cell(0.5, 29.0); lane.set_rightmost_cell(0.5, 29.0)
    net = RoadNetwork(30.0); net.add_lane(lane)
    with pytest.raises(AssertionError, match="CFL"):
        net.forward(0.01, True)
    net.forward(0.001, True)                          # flags were cleared; a valid step goes through
