"""Wall time per gradient-descent episode of configs 1-3 through the DROP-IN object API (road.* / RoadNetwork.forward,
one kernel launch per lane and step, the reference's own loop structure: example/inverse/_inverse.py:187-242)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch as th
import dhts_b200.dropin as dropin
from dhts_b200.dropin import runtime as rt
dropin.install(); rt.configure(precision="mixed")
from road.lane.dmacro_lane import dMacroLane
from road.lane.dmicro_lane import dMicroLane
from road.network.road_network import RoadNetwork
from road.network.route import MicroRoute
from road.vehicle.micro_vehicle import MicroVehicle

N, dx, umax, dt, T = 10, 5.0, 30.0, 0.01, 500
th.manual_seed(0); np.random.seed(0)


def clear(net):
    for lane in net.lane.values():
        lane.clear()
    net.vehicle.clear(); net.micro_route.clear(); net.num_vehicle = 0


def macro_or_hybrid(mode, episodes=4):
    net = RoadNetwork(umax)
    lane = dMacroLane(0, N * dx, umax, dx)
    lane.set_leftmost_cell(th.rand(()), th.rand(()) * umax); lane.set_rightmost_cell(th.rand(()), th.rand(()) * umax)
    net.add_lane(lane)
    if mode == "hybrid":
        net.add_lane(dMicroLane(1, N * dx, umax))
        l2 = dMacroLane(2, N * dx, umax, dx)
        l2.set_leftmost_cell(th.rand(()), th.rand(()) * umax); l2.set_rightmost_cell(th.rand(()), th.rand(()) * umax)
        net.add_lane(l2); net.connect_lane(0, 1); net.connect_lane(1, 2)
        net.macro_route = net.create_random_macro_route()
    er = th.rand(N).requires_grad_(); eu = (th.rand(N) * umax).requires_grad_()
    tr, tu = th.rand(N), th.rand(N) * umax
    opt = th.optim.Adam((er, eu), lr=1e-3)
    times = []
    for ep in range(episodes):
        th.cuda.synchronize(); t0 = time.time()
        clear(net)
        net.lane[0].set_state_vector_u(er, eu)
        for _ in range(T):
            net.forward(dt, True)
        s = net.lane[0].get_state_vector()
        err = th.pow(tr - s[0], 2.0).sum() + th.pow(tu - s[2], 2.0).sum()
        opt.zero_grad(); err.backward(); opt.step()
        th.cuda.synchronize(); times.append(time.time() - t0)
    return times


def micro(episodes=4):
    n = 10
    net = RoadNetwork(umax)
    net.add_lane(dMicroLane(0, 1e10, umax))
    ep_ = (th.arange(n) * 20.0 + th.rand(n) * 10.0).requires_grad_(); ev_ = (9.0 + th.rand(n) * 12.0).requires_grad_()
    tp, tv = th.arange(n) * 20.0 + 5.0, th.full((n,), 15.0)
    opt = th.optim.Adam((ep_, ev_), lr=1e-2)
    times = []
    for ep in range(episodes):
        th.cuda.synchronize(); t0 = time.time()
        clear(net)
        net.vehicle.clear(); net.micro_route.clear()
        for i in range(n):
            mv = MicroVehicle.default_micro_vehicle(umax); mv.position = ep_[i]; mv.speed = ev_[i]
            net.add_vehicle(mv, MicroRoute([0]))
        net.lane[0].set_state_vector(ep_, ev_)
        for _ in range(T):
            net.forward(dt, True)
        s = net.lane[0].get_state_vector()
        err = th.pow(tp - s[0], 2.0).sum() + th.pow(tv - s[1], 2.0).sum()
        opt.zero_grad(); err.backward(); opt.step()
        th.cuda.synchronize(); times.append(time.time() - t0)
    return times


for name, fn in (("macro  (config 1)", lambda: macro_or_hybrid("macro")), ("micro  (config 2)", micro),
                 ("hybrid (config 3)", lambda: macro_or_hybrid("hybrid"))):
    t = fn()
    print("%s: s/episode %s  (steady %.3f)" % (name, [round(x, 3) for x in t], min(t[1:])))
