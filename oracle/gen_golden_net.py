"""TEST INFRASTRUCTURE: freeze outputs of the LIVE reference's ITSCP network (macro mode) into tests/golden/.

Run in THIS container only (needs /root/reference):

    python oracle/gen_golden_net.py            # fp64 tier (dtype-proxied reference, SURVEY 8c tier 2)

The reference's ``ItscpRoadNetwork`` (example/control/itscp/_simulator.py) and its lanes are imported
UNMODIFIED; ``example/control/itscp/_env.py`` itself cannot be imported here (highway_env, gym and pygame are
absent), so the three pieces of it that drive the simulator -- the lane graph of ``_make_road`` (:225-439), the
per-frame loop of ``_simulate_step`` (:588-700) and ``lane_signal_info`` (:885-962) -- are driven from this
script: the lane graph comes from ``dhts_b200.itscp.ItscpGrid`` (checked below against the reference lanes it
builds), signals are computed with the reference's own ``dmath.sigmoid``, the reward loop calls the reference's
``RunningMean``.

Fixtures: tests/golden/itscp_macro_fp64.npz with cases `a_*` (2x2 grid, lanes start empty, as the env does)
and `b_*` (1 intersection x 2 lanes, random initial state): every state, reward, sigmoid constants, gradients
wrt action, per-frame lane signals, inflow, initial state.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DHTS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.path.insert(2, HERE)

from gen_golden import switch_fp64  # noqa: E402


def load_reference_simulator():
    """example/ is a namespace package shadowed by a site-packages example.py: load the two files by path."""
    for name in ("example", "example.common", "example.control", "example.control.itscp"):
        m = types.ModuleType(name); m.__path__ = []
        sys.modules[name] = m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec); sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    rms = load("example.common.rms", "example/common/rms.py")
    sim = load("example.control.itscp._simulator", "example/control/itscp/_simulator.py")
    return sim.ItscpRoadNetwork, rms.RunningMean


def f(x):
    return float(x.item()) if isinstance(x, th.Tensor) else float(x)


def run_case(tag, grid, T, frames_per_signal, seed, random_init, umax=60.0, freq=30, veh_len=5.0, static_speed=0.2):
    from dmath.operation import sigmoid
    from road.lane.dmacro_lane import dMacroLane
    ItscpRoadNetwork, RunningMean = load_reference_simulator()
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    dt = 1.0 / freq
    net = ItscpRoadNetwork(umax)
    for info in grid.lanes:
        net.add_lane(dMacroLane(len(net.lane), info.length, umax, grid.cell_length))
    for a, b in grid.links:
        net.connect_lane(a, b)
    L = grid.L
    for l in range(L):      # the restated lane graph must give the reference's discretisation
        assert net.lane[l].num_cell == grid.num_cell[l] and abs(net.lane[l].cell_length - grid.dx[l]) < 1e-12
    off = np.concatenate([[0], np.cumsum(grid.num_cell)]).astype(int)
    NC = int(off[-1])
    n2 = grid.num_intersection ** 2
    n_phase = max(1, T // frames_per_signal)
    action = th.tensor(rng.uniform(0.1, 0.9, n_phase * n2), requires_grad=True)
    bl = grid.boundary_lanes()
    # inflow schedule: sessions of constant density like itscp_random_schedule (_env.py:57-87)
    inc_np = np.zeros((T, L))
    for l in bl:
        for s0 in range(0, T, max(1, T // 5)):
            inc_np[s0:s0 + max(1, T // 5), l] = rng.uniform(0, 1)
    inc = {(t, l): th.tensor(inc_np[t, l], requires_grad=True) for t in range(T) for l in bl}
    routes = [net.create_random_macro_route() for _ in range(T)]
    route_tab = np.array([[[r.get_prev_lane(l) for l in range(L)], [r.get_next_lane(l) for l in range(L)]] for r in routes],
                         dtype=np.int32)
    # initial state
    if random_init:
        r0_np = rng.uniform(0.02, 0.9, NC); u0_np = rng.uniform(0.0, 1.0, NC) * umax * (1.0 - 0.7 * r0_np)
    else:
        r0_np = np.zeros(NC); u0_np = np.full(NC, umax)
    r0 = th.tensor(r0_np, requires_grad=True); u0 = th.tensor(u0_np, requires_grad=True)
    for l in range(L):
        net.lane[l].set_state_vector_u(r0[off[l]:off[l + 1]], u0[off[l]:off[l + 1]])
    rms = RunningMean(100_000)
    sig_t = {}
    hist = np.zeros((T + 1, 3, NC)); kconst = np.zeros((T, NC)); sig_np = np.ones((T, L))
    reward = 0

    def snapshot(t):
        for l in range(L):
            for i, c in enumerate(net.lane[l].curr_cell):
                hist[t, 0, off[l] + i] = f(c.state.q.r); hist[t, 1, off[l] + i] = f(c.state.q.y); hist[t, 2, off[l] + i] = f(c.state.u)

    snapshot(0)
    for t in range(T):
        # lane_signal_info (_env.py:885-962), next_signal only (what _simulate_step stores, :600-603)
        phase = min(t // frames_per_signal, n_phase - 1)
        progress = min((t % frames_per_signal) / frames_per_signal, 1.0)
        for l, info in enumerate(grid.lanes):
            if info.loc == "mid" or not info.approaching:
                s = 1.0
            else:
                a = action[phase * n2 + info.row * grid.num_intersection + info.col]
                s = sigmoid(a - progress, constant=32) if info.loc in ("west", "east") else sigmoid(progress - a, constant=32)
                s.retain_grad(); sig_t[(t, l)] = s
                sig_np[t, l] = f(s)
            net.lane_signal[l] = s
            net.lane_incoming[l] = inc[(t, l)] if l in bl else -1
        net.macro_route = routes[t]
        net.forward(dt, True)
        snapshot(t + 1)
        # queue length (_env.py:618-648 with _is_static_speed :557-575)
        for l in range(L):
            lane = net.lane[l]
            q = 0
            for i, c in enumerate(lane.curr_cell):
                speed = c.state.u if isinstance(c.state.u, th.Tensor) else th.tensor(c.state.u)
                with th.no_grad():
                    rms.update((static_speed - speed).cpu().numpy())
                    constant = 16.0 / np.abs(rms.mean())
                kconst[t, off[l] + i] = constant
                q = q + sigmoid(static_speed - speed, constant=constant) * (c.state.q.r * lane.cell_length / net.vehicle_length)
            reward = reward + (-1.0) * ((q ** 2.0) * dt)
    # terminal term so that every cell of the last state matters
    w_r = rng.normal(size=NC); w_u = rng.normal(size=NC) / umax
    term = 0
    for l in range(L):
        for i, c in enumerate(net.lane[l].curr_cell):
            term = term + w_r[off[l] + i] * c.state.q.r + w_u[off[l] + i] * c.state.u
    loss = reward + term
    loss.backward()
    g_sig = np.zeros((T, L)); g_inc = np.zeros((T, L))
    for (t, l), s in sig_t.items():
        g_sig[t, l] = 0.0 if s.grad is None else f(s.grad)
    for (t, l), x in inc.items():
        g_inc[t, l] = 0.0 if x.grad is None else f(x.grad)
    out = dict(T=T, frames_per_signal=frames_per_signal, umax=umax, dt=dt, veh_len=veh_len, static_speed=static_speed,
               num_intersection=grid.num_intersection, num_lane=grid.num_lane, lane_length=grid.lane_length,
               cell_length=grid.cell_length, action=action.detach().numpy(), incoming=inc_np, route=route_tab, r0=r0_np,
               u0=u0_np, hist=hist, kconst=kconst, sig=sig_np, reward=f(reward), term=f(term), w_r=w_r, w_u=w_u,
               g_action=action.grad.numpy(), g_sig=g_sig, g_inc=g_inc, g_r0=r0.grad.numpy(), g_u0=u0.grad.numpy())
    print(tag, "L", L, "NC", NC, "T", T, "reward %.6g term %.6g |g_action| %.3g |g_r0| %.3g max r %.3g" %
          (out["reward"], out["term"], np.abs(out["g_action"]).max(), np.abs(out["g_r0"]).max(), hist[:, 0].max()))
    return {tag + "_" + k: v for k, v in out.items()}


def main():
    switch_fp64()
    from dhts_b200.itscp import ItscpGrid
    out = {}
    out.update(run_case("a", ItscpGrid(2, 1, 10.0, 5.0), T=90, frames_per_signal=30, seed=11, random_init=False))
    out.update(run_case("b", ItscpGrid(1, 2, 12.0, 5.0), T=60, frames_per_signal=20, seed=12, random_init=True))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "itscp_macro_fp64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
