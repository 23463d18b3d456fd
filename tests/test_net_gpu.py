"""GPU parity of the fused connected-network rollout (dhts_net_rollout_{fwd,bwd}_*, through the C ABI) against
fixtures frozen from the live reference's ItscpRoadNetwork and against the CPU checker on seeded random networks.
fp64 tolerance asserted: 1e-9 of the largest entry (north-star bar: rtol 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from net_cases import fixture_case, grid_of, random_network, random_routes, reward_and_injection

pytestmark = pytest.mark.gpu
F64 = torch.float64


def t64(a, dev, grad=False):
    t = torch.tensor(np.asarray(a, dtype=np.float64), dtype=F64, device=dev)
    return t.requires_grad_() if grad else t


@pytest.mark.parametrize("tag", ["a", "b"])
def test_itscp_macro_matches_live_reference(dev, tag):
    """Whole pipeline a user calls: action -> lane signals -> fused rollout -> queue reward with the running-mean
    constants (+ a terminal term), gradients wrt action, inflow and initial state."""
    from dhts_b200.itscp import ItscpBatch
    G = fixture_case(tag)
    grid = grid_of(G)
    T = int(G["T"])
    env = ItscpBatch(grid, dev, speed_limit=float(G["umax"]), simulation_frequency=round(1.0 / float(G["dt"])),
                     signal_length=int(G["frames_per_signal"]) * float(G["dt"]), vehicle_length=float(G["veh_len"]),
                     static_speed=float(G["static_speed"]))
    assert env.frames_per_signal == int(G["frames_per_signal"])
    action = t64(G["action"][None], dev, True)
    inc = t64(G["incoming"][None], dev, True)
    r0 = t64(G["r0"][None], dev, True); u0 = t64(G["u0"][None], dev, True)
    route = torch.tensor(G["route"], dtype=torch.int32, device=dev)
    reward, states = env.rollout(action, inc, route, T, differentiable=True, r0=r0, u0=u0, exact_constants=True)
    hist = states[:, 0].detach().cpu().numpy()
    assert np.abs(hist - G["hist"]).max() < 1e-10
    assert abs(float(reward[0].detach()) - float(G["reward"])) < 1e-6 * abs(float(G["reward"]))      # float32-rounded constants in the reference
    term = (states[T, 0, 0] * t64(G["w_r"], dev)).sum() + (states[T, 0, 2] * t64(G["w_u"], dev)).sum()
    assert abs(float(term.detach()) - float(G["term"])) < 1e-9 * abs(float(G["term"]))
    (reward[0] + term).backward()
    # 1e-6: the reference's sigmoid constants are float32-rounded (2.7e-7 relative, tests/test_net_oracle.py)
    assert relerr(action.grad[0].cpu().numpy(), G["g_action"]) < 2e-6
    assert relerr(inc.grad[0].cpu().numpy(), G["g_inc"]) < 2e-6
    assert relerr(r0.grad[0].cpu().numpy(), G["g_r0"]) < 2e-6
    assert relerr(u0.grad[0].cpu().numpy(), G["g_u0"], floor=1e-9) < 2e-6


@pytest.mark.parametrize("tag", ["a", "b"])
def test_kernel_adjoint_with_reference_constants(dev, tag):
    """Same, but with the fixture's own constants injected, so that nothing but the kernels differs: 1e-9."""
    from dhts_b200 import Flags
    from dhts_b200.network import net_rollout
    G = fixture_case(tag)
    grid = grid_of(G)
    topo = grid.topology(dev)
    T, umax, dt = int(G["T"]), float(G["umax"]), float(G["dt"])
    off = np.array(topo.cell_off)
    _, gst = reward_and_injection(G["hist"], off, topo.cell_length, G["kconst"], dt, float(G["veh_len"]), float(G["static_speed"]))
    gst[T - 1, 0] += G["w_r"]; gst[T - 1, 2] += G["w_u"]
    sig = t64(G["sig"][None], dev, True); inc = t64(G["incoming"][None], dev, True)
    r0 = t64(G["r0"][None], dev, True); u0 = t64(G["u0"][None], dev, True)
    flags = Flags(dev)
    states, _ = net_rollout(topo, r0, u0, umax, dt, T, sig=sig, incoming=inc, route=torch.tensor(G["route"], dtype=torch.int32, device=dev),
                            soft=True, flags=flags)
    (states[1:, 0] * t64(gst, dev)).sum().backward()
    flags.check()
    assert np.abs(states[:, 0].detach().cpu().numpy() - G["hist"]).max() < 1e-10
    tensor_sig = np.array([(i.loc != "mid" and i.approaching) for i in grid.lanes])
    assert relerr(sig.grad[0].cpu().numpy()[:, tensor_sig], G["g_sig"][:, tensor_sig]) < 1e-9
    assert relerr(inc.grad[0].cpu().numpy(), G["g_inc"]) < 1e-7
    assert relerr(r0.grad[0].cpu().numpy(), G["g_r0"]) < 1e-7
    assert relerr(u0.grad[0].cpu().numpy(), G["g_u0"], floor=1e-9) < 1e-7


@pytest.mark.parametrize("mode,seed", [(0, 1), (0, 2), (1, 3), (1, 4)])
def test_random_networks_match_checker(dev, mode, seed):
    """Random lane graphs (0, 1, several neighbours per side; 1-4 cells per lane; own ghosts with non-default
    records), 3 replicas with different inputs, per-replica route schedules, fused queue reward (qk) and an injected
    per-step adjoint: kernels vs the CPU checker."""
    from dhts_b200 import Flags
    from dhts_b200.network import MacroNetTopology, net_rollout
    from oracle import net_oracle as NO
    rng = np.random.default_rng(100 + seed)
    L, T, R, umax, dt = 23, 40, 3, 30.0, 0.02
    num_cell, dx, links = random_network(rng, L)
    topo = MacroNetTopology(num_cell, dx, links, dev, mode)
    net = NO.Net(num_cell, dx, links, mode)
    assert net.n_own == topo.n_own and net.NC == topo.NC
    NC = topo.NC
    r0 = rng.uniform(0.02, 0.9, (R, NC)); u0 = rng.uniform(0, 1, (R, NC)) * umax * (1 - 0.6 * r0)
    own0 = np.stack([rng.uniform(0, 1, (R, max(topo.n_own, 1))), rng.uniform(0, umax, (R, max(topo.n_own, 1)))], -1)[:, :topo.n_own]
    sig = rng.uniform(0.2, 0.8, (R, T, L)); inc = rng.uniform(0, 1, (R, T, L))
    route = np.stack([random_routes(rng, net.prev, net.next, T) for _ in range(R)])
    qk = rng.uniform(0.5, 3.0, T)
    gst = rng.normal(size=(R, T, 3, NC)) * 0.1
    g_rew = rng.normal(size=R)
    tr0, tu0 = t64(r0, dev, True), t64(u0, dev, True)
    tsig, tinc = t64(sig, dev, mode == 1), t64(inc, dev, mode == 1)
    town = t64(own0, dev, True) if topo.n_own else None
    flags = Flags(dev)
    states, reward = net_rollout(topo, tr0, tu0, umax, dt, T, sig=tsig if mode else None, incoming=tinc if mode else None,
                                 route=torch.tensor(route, dtype=torch.int32, device=dev), own0=town, soft=True,
                                 qk=t64(qk, dev), veh_len=5.0, static_speed=8.0, flags=flags)
    loss = (states[1:].permute(1, 0, 2, 3) * t64(gst, dev)).sum() + (reward * t64(g_rew, dev)).sum()
    loss.backward()
    flags.check()
    for b in range(R):
        o = NO.rollout(net, r0[b], u0[b], umax, dt, T, sig=sig[b] if mode else None, incoming=inc[b] if mode else None,
                       route=route[b], own0=own0[b] if topo.n_own else None, soft=True, qk=qk, veh_len=5.0, static_speed=8.0,
                       g_states=gst[b], g_reward=g_rew[b], want_grad=True)
        assert o["cfl"] == 0
        assert np.abs(states[:, b].detach().cpu().numpy() - o["hist"]).max() < 1e-10
        assert abs(float(reward[b]) - o["reward"]) < 1e-10 * max(1.0, abs(o["reward"]))
        assert relerr(tr0.grad[b].cpu().numpy(), o["g_r0"]) < 1e-9
        assert relerr(tu0.grad[b].cpu().numpy(), o["g_u0"]) < 1e-9
        if topo.n_own:
            assert relerr(town.grad[b].cpu().numpy(), o["g_own0"][:topo.n_own]) < 1e-9
        if mode == 1:
            assert relerr(tsig.grad[b].cpu().numpy(), o["g_sig"]) < 1e-9
            assert relerr(tinc.grad[b].cpu().numpy(), o["g_inc"]) < 1e-9


def test_replicas_are_independent_and_hard_signals(dev):
    """Permuting replicas permutes results bitwise; the non-differentiable signal test (sig > 0.5) matches the checker."""
    from dhts_b200.network import net_rollout
    from oracle import net_oracle as NO
    G = fixture_case("b")
    grid = grid_of(G)
    topo = grid.topology(dev)
    T, umax, dt = 30, float(G["umax"]), float(G["dt"])
    rng = np.random.default_rng(5)
    R = 6
    NC, L = topo.NC, topo.L
    r0 = rng.uniform(0.02, 0.9, (R, NC)); u0 = rng.uniform(0, 1, (R, NC)) * umax * (1 - 0.7 * r0)
    sig = (rng.uniform(0, 1, (R, T, L)) > 0.4).astype(np.float64); inc = rng.uniform(0, 1, (R, T, L))
    route = torch.tensor(G["route"][:T], dtype=torch.int32, device=dev)
    run = lambda p: net_rollout(topo, t64(r0[p], dev), t64(u0[p], dev), umax, dt, T, sig=t64(sig[p], dev), incoming=t64(inc[p], dev),
                                route=route, soft=False)[0]
    perm = rng.permutation(R)
    a, b = run(np.arange(R)), run(perm)
    assert torch.equal(a[:, perm], b)
    net = NO.Net(grid.num_cell, grid.dx, grid.links, 1)
    o = NO.rollout(net, r0[0], u0[0], umax, dt, T, sig=sig[0], incoming=inc[0], route=G["route"][:T], soft=False)
    assert np.abs(a[:, 0].cpu().numpy() - o["hist"]).max() < 1e-10


def test_missing_route_raises_like_the_reference(dev):
    """A lane with two successors and no route entry: the reference indexes self.lane[-1] (KeyError)."""
    from dhts_b200 import Flags
    from dhts_b200.network import MacroNetTopology, net_rollout
    topo = MacroNetTopology([2, 2, 2], [5.0, 5.0, 5.0], [(0, 1), (0, 2)], dev, 0)
    r0 = torch.full((1, 6), 0.3, dtype=F64, device=dev); u0 = torch.full((1, 6), 10.0, dtype=F64, device=dev)
    flags = Flags(dev)
    net_rollout(topo, r0, u0, 30.0, 0.01, 2, route=torch.full((2, 2, 3), -1, dtype=torch.int32, device=dev), flags=flags)
    with pytest.raises(KeyError):
        flags.check()


def test_f32_build_tracks_f64(dev):
    from dhts_b200.network import net_rollout
    G = fixture_case("b")
    grid = grid_of(G)
    topo = grid.topology(dev)
    T, umax, dt = int(G["T"]), float(G["umax"]), float(G["dt"])
    f32 = lambda a: torch.tensor(a[None], dtype=torch.float32, device=dev)
    st, _ = net_rollout(topo, f32(G["r0"]), f32(G["u0"]), umax, dt, T, sig=f32(G["sig"]), incoming=f32(G["incoming"]),
                        route=torch.tensor(G["route"], dtype=torch.int32, device=dev), soft=True)
    h = st[:, 0].cpu().numpy().astype(np.float64)
    assert np.isfinite(h).all()
    assert np.abs(h[:, 0] - G["hist"][:, 0]).max() < 5e-3      # fp32 storage and arithmetic over 60 coupled steps
