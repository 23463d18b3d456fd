"""GPU parity of the fused HYBRID-network rollout (dhts_hyb_rollout_{fwd,bwd}_*, through the C ABI) against fixtures
frozen from the live reference's ItscpRoadNetwork in hybrid mode (oracle/gen_golden_hyb.py: 3x3 grid, lanes of the
centre intersection are dMicroLanes, everything else dMacroLanes; signal-blended ghosts and head deltas, cross-lane
leaders, every Conversion.* in lane-id order).  fp64 tolerance asserted: states 1e-9 absolute, gradients 1e-8 of the
largest entry (north-star bar: rtol 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from hyb_cases import build, fixture_case, run_fixture, spawn_routes, t64

pytestmark = pytest.mark.gpu
F64 = torch.float64


@pytest.mark.parametrize("tag", ["h", "g"])
def test_hybrid_itscp_matches_live_reference(dev, tag):
    """action -> signals -> fused hybrid rollout -> queue reward + terminal term; every stored quantity of every frame
    (cells, vehicle counts, vehicles head first, head deltas) and the gradients wrt action / signals / inflow /
    initial state against the live reference."""
    G = fixture_case(tag)
    o = run_fixture(G, dev)
    st, topo = o["st"], o["topo"]
    T = int(G["T"])
    o["flags"].check()
    cells = st.cells[:, 0].detach().cpu().numpy()
    assert np.abs(cells - G["hist"]).max() < 1e-9
    cnt = st.count[:, 0].cpu().numpy()
    assert (cnt == G["vcnt"]).all()
    assert int(G["vid"].max()) + 1 >= 4, "fixture must exercise spawns"
    p, v, a = (x[:, 0].detach().cpu().numpy() for x in (o["p"], o["v"], o["a"]))
    ours = np.stack([p, v, a], -1)                                   # [T+1, ML, cap, 3] head first
    mask = (np.arange(ours.shape[2])[None, None] < G["vcnt"][..., None])
    assert np.abs((ours - G["veh"])[mask]).max() < 1e-9
    head = st.head[:, 0].detach().cpu().numpy()
    assert np.abs(head - G["head"]).max() < 1e-7 * max(1.0, np.abs(G["head"]).max())
    assert abs(float(o["reward"].detach()) - float(G["reward"])) < 1e-9 * abs(float(G["reward"]))
    assert abs(float(o["term"].detach()) - float(G["term"])) < 1e-9 * abs(float(G["term"]))
    (o["reward"] + o["term"]).backward()
    o["flags"].check()
    lanes = [l for l, info in enumerate(o["grid"].lanes) if info.loc != "mid" and info.approaching]
    assert relerr(o["action"].grad[0].cpu().numpy(), G["g_action"]) < 1e-8
    assert relerr(o["sig"].grad[0].cpu().numpy()[:, lanes], G["g_sig"][:, lanes]) < 1e-8
    assert relerr(o["inc"].grad[0].cpu().numpy(), G["g_inc"]) < 1e-8
    assert relerr(o["r0"].grad[0].cpu().numpy(), G["g_r0"]) < 1e-8
    assert relerr(o["u0"].grad[0].cpu().numpy(), G["g_u0"], floor=1e-9) < 1e-8


def test_hybrid_replicas_are_independent(dev):
    """R replicas with permuted inputs give bitwise permuted trajectories and gradients (one CTA per replica, no
    cross-replica state)."""
    from dhts_b200 import Flags
    from dhts_b200.hybrid_network import hybrid_rollout
    G = fixture_case("h")
    grid, topo = build(G, dev)
    T, umax, dt = 60, float(G["umax"]), float(G["dt"])
    rng = np.random.default_rng(11)
    R = 5
    # gentle perturbations of the fixture: large ones drive cells to vacuum and vehicles into collisions, where the
    # reference's IDM Jacobian divides by the raw gap and the NaN-gradient flag is raised
    r0 = np.clip(G["r0"][None] * rng.uniform(0.95, 1.05, (R, topo.NC)), 0.01, 0.95)
    u0 = G["u0"][None] * rng.uniform(0.95, 1.0, (R, topo.NC))
    sig = np.clip(G["sig"][None, :T] + rng.uniform(-0.05, 0.05, (R, T, topo.L)), 0.0, 1.0)
    inc = np.clip(G["incoming"][None, :T] * rng.uniform(0.9, 1.1, (R, T, topo.L)), 0.0, 1.0)
    route = torch.tensor(G["route"][:T], dtype=torch.int32, device=dev)
    sp = torch.tensor(spawn_routes(G, topo), dtype=torch.int32, device=dev)
    w = t64(rng.normal(size=(topo.NC,)), dev)

    def run(perm):
        tr, tu = t64(r0[perm], dev, True), t64(u0[perm], dev, True)
        ts = t64(sig[perm], dev, True)
        flags = Flags(dev)
        st = hybrid_rollout(topo, tr, tu, umax, dt, T, sig=ts, incoming=t64(inc[perm], dev), route=route, spawn_route=sp,
                            flags=flags)
        ((st.cells[T, :, 0] * w).sum() + st.speed[T].sum()).backward()
        assert (flags.read()[0] & ~4) == 0 and bool(torch.isfinite(tr.grad).all()) and bool(torch.isfinite(ts.grad).all())
        return st.hist.detach(), st.aux.detach(), tr.grad, ts.grad

    perm = rng.permutation(R)
    a, b = run(np.arange(R)), run(perm)
    assert torch.equal(a[0][:, perm], b[0]) and torch.equal(a[1][:, perm], b[1])
    assert torch.equal(a[2][perm], b[2]) and torch.equal(a[3][perm], b[3])
    assert float(a[1][..., topo.A_CNT:topo.A_CNT + topo.ML].max()) >= 1, "vehicles must have been spawned"


def test_hybrid_collisions_and_nan_gradients_are_flagged(dev):
    """Inputs far from the fixture (speeds scaled by up to 1.3) drive vehicles into collisions: the forward pass counts
    them (print-and-continue, _micro_lane.py:151-162) and a non-finite gradient raises the reference's NaN assert
    (dmacro_lane.py:308) instead of passing silently."""
    from dhts_b200 import Flags
    from dhts_b200._lib import FLAG_COLLISION, FLAG_NAN_GRAD
    from dhts_b200.hybrid_network import hybrid_rollout
    G = fixture_case("h")
    grid, topo = build(G, dev)
    T, umax, dt = 60, float(G["umax"]), float(G["dt"])
    rng = np.random.default_rng(11)
    R = 5
    r0 = np.clip(G["r0"][None] * rng.uniform(0.7, 1.3, (R, topo.NC)), 0.01, 0.95)
    u0 = G["u0"][None] * rng.uniform(0.7, 1.3, (R, topo.NC))
    sig = np.clip(G["sig"][None, :T] + rng.uniform(-0.2, 0.2, (R, T, topo.L)), 0.0, 1.0)
    inc = np.clip(G["incoming"][None, :T] * rng.uniform(0.5, 1.5, (R, T, topo.L)), 0.0, 1.0)
    tr = t64(r0, dev, True)
    flags = Flags(dev)
    st = hybrid_rollout(topo, tr, t64(u0, dev), umax, dt, T, sig=t64(sig, dev), incoming=t64(inc, dev),
                        route=torch.tensor(G["route"][:T], dtype=torch.int32, device=dev),
                        spawn_route=torch.tensor(spawn_routes(G, topo), dtype=torch.int32, device=dev), flags=flags)
    assert bool(torch.isfinite(st.hist).all())
    bits, ncol = flags.read()
    assert (bits & FLAG_COLLISION) and ncol > 0
    st.speed[T].sum().backward()
    finite = bool(torch.isfinite(tr.grad).all())
    bits, _ = flags.read()
    assert finite == (not (bits & FLAG_NAN_GRAD))
    if not finite:
        with pytest.raises(AssertionError):
            flags.check(quiet_collisions=True)


def test_plain_mode_chain_matches_live_reference(dev):
    """MODE_PLAIN (no signals): the macro(10) -> micro -> macro(10) chain of SURVEY A.5 / config 3 through the FUSED
    hybrid rollout against tests/golden/hybrid_chain_fp64.npz (live reference, 700 steps, 19 spawns, 12 absorptions):
    spawn / on-lane counts and capacitor per step, final cells and vehicles, loss, gradients reaching lane 0 from a loss
    on lane 2."""
    from conftest import golden
    from dhts_b200 import Flags
    from dhts_b200.hybrid_network import HybridNetTopology, hybrid_rollout
    from dhts_b200.network import MODE_PLAIN
    g = golden("hybrid_chain_fp64")
    N, dx, umax, dt, T = int(g["N"]), float(g["dx"]), float(g["umax"]), float(g["dt"]), int(g["T"])
    topo = HybridNetTopology([0, 1, 0], [N, 0, N], [dx, 1.0, dx], [N * dx] * 3, [(0, 1), (1, 2)], dev, MODE_PLAIN, veh_cap=12)
    assert topo.n_own == 4 and topo.NCAP == 1 and topo.routes == [(1, 2)]
    gh = g["ghost_ru"]
    own0 = t64(np.stack([gh[0], gh[2], gh[1], gh[3]])[None], dev)        # own slots: (left l0, left l2, right l0, right l2)
    r0 = t64(g["r0"].reshape(1, -1), dev, True); u0 = t64(g["u0"].reshape(1, -1), dev, True)
    route = torch.tensor([[[-1, 0, -1], [1, -1, -1]]] * T, dtype=torch.int32, device=dev)     # create_random_macro_route: 0 -> 1 only
    sp = torch.zeros((1, 32), dtype=torch.int32, device=dev)
    flags = Flags(dev)
    st = hybrid_rollout(topo, r0, u0, umax, dt, T, route=route, spawn_route=sp, own0=own0, flags=flags)
    flags.check()
    cnt = st.count[1:, 0, 0].cpu().numpy()
    assert (cnt == g["nveh_hist"]).all()
    nsp = st.aux[1:, 0, topo.A_NSP].detach().round().long().cpu().numpy()
    assert (nsp == g["nspawn_hist"]).all() and nsp[-1] == 19
    assert np.abs(st.capacitor[1:, 0, 0].detach().cpu().numpy() - g["cap_hist"]).max() < 1e-9
    cells = st.cells[T, 0].detach().cpu().numpy()                          # [3, 20]
    assert np.abs(cells[:, :N] - g["lane0"]).max() < 1e-9 and np.abs(cells[:, N:] - g["lane2"]).max() < 1e-9
    p, v, a, valid = st.by_rank()
    n = int(cnt[-1])
    tail_first = lambda x: torch.flip(x[T, 0, 0, :n], dims=[0])
    veh = torch.stack([tail_first(p), tail_first(v), tail_first(a)], -1)
    assert veh.shape == g["veh"].shape and np.abs(veh.detach().cpu().numpy() - g["veh"]).max() < 1e-9
    loss = (st.cells[T, 0, 0, N:] * t64(g["w_r"], dev)).sum() + (st.cells[T, 0, 2, N:] * t64(g["w_u"], dev)).sum()
    w = g["w_veh"]
    for i in range(n):
        loss = loss + float(w[2 * i % 8]) * veh[i, 0] * 0.01 + float(w[(2 * i + 1) % 8]) * veh[i, 1] * 0.01
    assert abs(float(loss) - float(g["loss"])) < 1e-9
    loss.backward()
    flags.check()
    gr, gu = r0.grad[0].cpu().numpy(), u0.grad[0].cpu().numpy()
    assert relerr(gr[:N], g["g_r0_lane0"]) < 1e-8 and relerr(gu[:N], g["g_u0_lane0"]) < 1e-8
    assert relerr(gr[N:], g["g_r0_lane2"]) < 1e-8 and relerr(gu[N:], g["g_u0_lane2"]) < 1e-8
    assert np.abs(g["g_r0_lane0"]).max() > 0


def test_f32_build_tracks_f64_on_a_short_horizon(dev):
    """fp32 symbols of the hybrid rollout (tolerance stated separately from fp64): over 40 frames of fixture `h`
    the discrete events agree and the cell states stay within 2e-4 of the fp64 run (largest entry); gradients finite."""
    from dhts_b200 import Flags
    from dhts_b200.hybrid_network import hybrid_rollout
    G = fixture_case("h")
    grid, topo = build(G, dev)
    T, umax, dt = 40, float(G["umax"]), float(G["dt"])
    route = torch.tensor(G["route"][:T], dtype=torch.int32, device=dev)
    sp = torch.tensor(spawn_routes(G, topo), dtype=torch.int32, device=dev)
    out = {}
    for dt_t in (torch.float64, torch.float32):
        c = lambda a: torch.tensor(np.asarray(a), dtype=dt_t, device=dev)
        r0 = c(G["r0"][None]).requires_grad_(); u0 = c(G["u0"][None])
        flags = Flags(dev)
        st = hybrid_rollout(topo, r0, u0, umax, dt, T, sig=c(G["sig"][None, :T]), incoming=c(G["incoming"][None, :T]), route=route,
                            spawn_route=sp, flags=flags)
        (st.cells[T, :, 0].sum() + st.speed[T].sum()).backward()
        flags.check(quiet_collisions=True)
        out[dt_t] = (st.cells.detach().double().cpu().numpy(), st.count.cpu().numpy(), r0.grad.double().cpu().numpy())
    a, b = out[torch.float64], out[torch.float32]
    assert (a[1] == b[1]).all()
    assert np.abs(a[0] - b[0]).max() < 2e-4 * np.abs(a[0]).max()
    assert np.isfinite(b[2]).all() and relerr(b[2], a[2]) < 5e-2


def test_per_vehicle_idm_parameters_match_live_reference(dev):
    """Every vehicle with its OWN IDM parameter set (MicroVehicle.random_micro_vehicle, road/vehicle/micro_vehicle.py:74-122):
    macro(10) -> micro -> micro -> macro(10) against tests/golden/hybrid_chain_pv_fp64.npz (live reference, 600 steps; five
    heterogeneous initial vehicles that are handed from micro lane to micro lane and absorbed, nine default vehicles
    spawned behind them).  Counts per step, final cells and vehicles, loss, gradients wrt lane 0 and wrt the initial
    vehicle positions and speeds."""
    from conftest import golden
    from dhts_b200 import Flags
    from dhts_b200.hybrid_network import HybridNetTopology, default_vehicle_params, hybrid_rollout
    from dhts_b200.network import MODE_PLAIN
    g = golden("hybrid_chain_pv_fp64")
    N, dx, umax, dt, T = int(g["N"]), float(g["dx"]), float(g["umax"]), float(g["dt"]), int(g["T"])
    cap = 12
    topo = HybridNetTopology([0, 1, 1, 0], [N, 0, 0, N], [dx, 1.0, 1.0, dx], [N * dx] * 4, [(0, 1), (1, 2), (2, 3)], dev, MODE_PLAIN,
                             veh_cap=cap)
    assert topo.n_own == 4 and topo.NCAP == 1 and topo.ML == 2
    gh = g["ghost_ru"]
    own0 = t64(np.stack([gh[0], gh[2], gh[1], gh[3]])[None], dev)
    r0 = t64(g["r0"].reshape(1, -1), dev, True); u0 = t64(g["u0"].reshape(1, -1), dev, True)
    route = torch.tensor([[[-1, 0, -1, -1], [1, -1, -1, -1]]] * T, dtype=torch.int32, device=dev)
    sp = torch.full((2, 32), topo.route_id([1, 2, 3]), dtype=torch.int32, device=dev)
    n1, n2 = len(g["pos1"]), len(g["pos2"])
    pad = lambda x: np.concatenate([x, np.zeros(cap - len(x))])
    p0 = t64(np.stack([pad(g["pos1"]), pad(g["pos2"])])[None], dev, True)
    v0 = t64(np.stack([pad(g["spd1"]), pad(g["spd2"])])[None], dev, True)
    a0 = torch.full((1, 2, cap), 5.0, dtype=F64, device=dev)
    route0 = np.zeros((2, cap)); route0[0, :] = topo.route_id([1, 2, 3]); route0[1, :] = topo.route_id([2, 3])
    pid0 = np.zeros((2, cap)); pid0[0, :n1] = 1 + np.arange(n1); pid0[1, :n2] = 1 + n1 + np.arange(n2)
    par = np.concatenate([[default_vehicle_params(umax)], g["par1"], g["par2"]])
    aux0 = topo.make_aux0(1, F64, p0, v0, a0, route0, [n1, n2], pid0=pid0)
    flags = Flags(dev)
    st = hybrid_rollout(topo, r0, u0, umax, dt, T, route=route, spawn_route=sp, own0=own0, aux0=aux0, veh_params=par, flags=flags)
    flags.check()
    cnt = st.count[1:, 0].cpu().numpy()
    assert (cnt == g["cnt_hist"]).all()
    nsp = st.aux[1:, 0, topo.A_NSP].detach().round().long().cpu().numpy() + n1 + n2
    assert (nsp == g["nspawn_hist"]).all() and nsp[-1] == 14
    cells = st.cells[T, 0].detach().cpu().numpy()
    assert np.abs(cells[:, :N] - g["lane0"]).max() < 1e-9 and np.abs(cells[:, N:] - g["lane3"]).max() < 1e-9
    p, v, a, valid = st.by_rank()
    w = g["w_veh"]
    loss = (st.cells[T, 0, 0, N:] * t64(g["w_r"], dev)).sum() + (st.cells[T, 0, 2, N:] * t64(g["w_u"], dev)).sum()
    for m, key in ((0, "veh1"), (1, "veh2")):
        n = int(cnt[-1, m])
        tail_first = lambda x: torch.flip(x[T, 0, m, :n], dims=[0])
        veh = torch.stack([tail_first(p), tail_first(v), tail_first(a)], -1)
        assert veh.shape[0] == g[key].shape[0] and np.abs(veh.detach().cpu().numpy() - g[key][:, :3]).max() < 1e-9
        for i in range(n):
            loss = loss + float(w[2 * i % 8]) * veh[i, 0] * 0.01 + float(w[(2 * i + 1) % 8]) * veh[i, 1] * 0.01
    assert abs(float(loss) - float(g["loss"])) < 1e-9
    loss.backward()
    flags.check()
    gr, gu = r0.grad[0].cpu().numpy(), u0.grad[0].cpu().numpy()
    assert relerr(gr[:N], g["g_r0_lane0"]) < 1e-8 and relerr(gu[:N], g["g_u0_lane0"]) < 1e-8
    assert relerr(gr[N:], g["g_r0_lane3"]) < 1e-8 and relerr(gu[N:], g["g_u0_lane3"]) < 1e-8
    gp, gv = p0.grad[0].cpu().numpy(), v0.grad[0].cpu().numpy()
    assert relerr(gp[0, :n1], g["g_pos1"]) < 1e-8 and relerr(gv[0, :n1], g["g_spd1"]) < 1e-8
    assert relerr(gp[1, :n2], g["g_pos2"]) < 1e-8 and relerr(gv[1, :n2], g["g_spd2"]) < 1e-8
    assert np.abs(g["g_pos1"]).max() > 0
    # one parameter set for everybody is a different trajectory (the parameters matter)
    st1 = hybrid_rollout(topo, r0.detach(), u0.detach(), umax, dt, T, route=route, spawn_route=sp, own0=own0,
                         aux0=topo.make_aux0(1, F64, p0.detach(), v0.detach(), a0, route0, [n1, n2]), flags=Flags(dev))
    assert not (st1.count[1:, 0].cpu().numpy() == g["cnt_hist"]).all()
