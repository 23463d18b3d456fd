"""Deferred stepping of a CONNECTED drop-in network: the queued ``RoadNetwork.forward`` calls run as one
``dhts_hyb_rollout_fwd`` launch with R = 1 (plain mode), and the lane / vehicle objects are rebuilt from its result.

What the kernel does per step is the whole of road/network/road_network.py:79-173 (ghost cells from neighbouring
macro lanes, head-vehicle leaders along vehicle routes, the lane steps, every ``Conversion.*`` in lane-id order); what
is left to this file is object bookkeeping the reference does on the side:

  * vehicle identity -- vehicles enter a micro lane at the tail (spawned by ``macro_to_micro``) and leave at the
    head (``micro_to_macro`` / ``micro_to_none``), so after the queue the lane holds the last ``count`` of
    (vehicles it had, head first) + (vehicles spawned, in order).  Micro -> micro hand-offs would need a per-vehicle
    identity in the kernel; networks with a micro lane feeding another micro lane are stepped immediately;
  * ids, ``network.vehicle``, ``network.micro_route`` -- one entry per spawn in the reference's order (time, then
    lane id of the spawning macro lane, road_network.py:113-127), routes drawn by REPLAYING
    ``create_random_route`` so that ``np.random`` ends in the state the reference leaves it in.  Routes must not
    depend on that state: every micro lane has at most one successor (else: immediate stepping);
  * parameters -- the kernel steps one IDM parameter set, so every vehicle present must be a
    ``default_micro_vehicle`` of the network's speed limit (what ``macro_to_micro`` spawns, conversion.py:44).

Any precondition that fails, and any overflow the kernel reports, replays the queue step by step instead; nothing
is changed before the launch has been validated.
"""
from __future__ import annotations

import numpy as np
import torch

from dhts_b200 import _lib
from dhts_b200.dropin import runtime as rt
from dhts_b200.hybrid_network import HybRolloutFn, HybridNetTopology, default_vehicle_params
from dhts_b200.network import MODE_PLAIN

MAX_CAP = 96


def _topology(net, dev):
    lanes = net.lane
    ids = sorted(lanes.keys())
    if ids != list(range(len(ids))):
        return None
    sig = (tuple((lanes[i].is_macro(), getattr(lanes[i], "num_cell", 0), float(getattr(lanes[i], "cell_length", 1.0)),
                  float(lanes[i].length), tuple(lanes[i].next_lane.keys())) for i in ids), float(net.vehicle_length), str(dev))
    hit = net._defer_cache.get("hyb_topo")
    if hit is not None and hit[0] == sig:
        return hit[1]
    for i in ids:
        l = lanes[i]
        if l.is_micro():
            if len(l.next_lane) > 1 or any(n.is_micro() for n in l.next_lane.values()):
                net._defer_cache["hyb_topo"] = (sig, None)
                return None
    longest = max([float(lanes[i].length) for i in ids if lanes[i].is_micro()] or [0.0])
    cap = int(longest / float(net.vehicle_length)) + 6
    topo = None
    if cap <= MAX_CAP:
        topo = HybridNetTopology.from_network(net, dev, MODE_PLAIN, veh_cap=cap)
    net._defer_cache["hyb_topo"] = (sig, topo)
    return topo


def flush_hyb(net, steps, dt, diff, replay):
    from model.macro._arz import ARZ
    from road.network.route import MicroRoute
    from road.vehicle.micro_vehicle import MicroVehicle
    sd, st = rt.step_dtype(), rt.store_dtype()
    dev, flags = rt.device(), rt.flags()
    topo = _topology(net, dev)
    if topo is None:
        return replay()
    lanes = net.lane
    L, cap, ML = topo.L, topo.veh_cap, topo.ML
    umax = float(net.speed_limit)
    par = default_vehicle_params(umax, float(net.vehicle_length))
    # ---- vehicles present: their own IDM parameter sets (set 0 = what spawned vehicles get), routes the topology knows,
    # head first per micro lane
    p0 = torch.zeros((1, max(ML, 1), cap), dtype=sd, device=dev); v0 = torch.zeros_like(p0); a0 = torch.zeros_like(p0)
    route0 = np.zeros((max(ML, 1), cap)); count0 = np.zeros(max(ML, 1)); pid0 = np.zeros((max(ML, 1), cap))
    par_sets = [tuple(float(x) for x in par)]
    par_idx = {par_sets[0]: 0}
    olds = []
    for m, l in enumerate(topo.micro):
        veh = lanes[l]._curr_vehicle
        if len(veh) > cap - 2:
            return replay()
        for mv in veh:
            if float(mv.idm_params()[5]) != par_sets[0][5] or mv.id not in net._micro_route:
                return replay()
        olds.append(list(reversed(veh)))
        if veh:
            n = len(veh)
            pp, vv = lanes[l]._state()
            p0[0, m, :n] = torch.flip(pp.to(sd), dims=[0]); v0[0, m, :n] = torch.flip(vv.to(sd), dims=[0])
            a0[0, m, :n] = rt.gather([mv.a for mv in olds[-1]], sd)
            count0[m] = n
            for k, mv in enumerate(olds[-1]):
                key = tuple(float(x) for x in mv.idm_params())
                if key not in par_idx:
                    par_idx[key] = len(par_sets); par_sets.append(key)
                pid0[m, k] = par_idx[key]
                mr = net._micro_route[mv.id]
                try:
                    route0[m, k] = topo.route_id(mr.route[mr.curr_idx:])
                except KeyError:
                    return replay()
    # ---- macro cells, own ghost records, capacitors
    macro = [l for l in range(L) if not topo.kind[l]]
    cat = lambda k: torch.cat([lanes[l]._vec("curr", k) for l in macro]).to(sd).reshape(1, -1)
    r0, y0, u0, e0 = cat("r"), cat("y"), cat("u").detach(), cat("e").detach()
    own = []
    for side, cell_of in ((0, lambda ln: ln.leftmost_cell), (1, lambda ln: ln.rightmost_cell)):
        for l in range(L):
            if topo.own_slot[side * L + l] >= 0:
                c = cell_of(lanes[l])
                own.append(torch.cat([lanes[l]._ghost(c, "r"), lanes[l]._ghost(c, "u")]))
    own0 = torch.stack(own).to(sd).reshape(1, topo.n_own, 2) if own else None
    cap_pairs = [(l, topo.cap_lane[j]) for l in range(L) for j in range(topo.cap_off[l], topo.cap_off[l + 1])]
    capac = rt.gather([lanes[a]._flux_capacitor.get(b, 0.0) for a, b in cap_pairs], sd).reshape(1, -1) if cap_pairs else None
    aux0 = topo.make_aux0(1, sd, p0[:, :ML], v0[:, :ML], a0[:, :ML], route0[:ML], count0[:ML], capac, pid0=pid0[:ML])
    par = [list(k) for k in par_sets]
    mroute = net._macro_route
    row = torch.tensor([[[mroute.get_prev_lane(l) for l in range(L)], [mroute.get_next_lane(l) for l in range(L)]]],
                       dtype=torch.int32, device=dev)
    route = row.expand(int(steps), 2, L).contiguous()
    # every spawn into micro lane m takes the one route that starts there
    n_pred = max([sum(1 for a, b in cap_pairs if b == l) for l in topo.micro] or [1])
    KS = max(1, int(steps) * max(1, n_pred))
    sp = torch.tensor([[topo.route_id(_walk(topo, l))] for l in topo.micro] or [[0]], dtype=torch.int32,
                      device=dev).expand(max(ML, 1), KS).contiguous()
    hist, auxh, headh = HybRolloutFn.apply(r0, y0, u0, own0, None, None, aux0, topo, route, sp, par, umax, float(dt), int(steps),
                                           True, flags.t, e0)
    bits, _ = flags.read()
    if bits & (_lib.FLAG_VEH_OVERFLOW | _lib.FLAG_ROUTE):
        flags.reset()
        return replay()
    # ---- unpack: cells
    last = hist[int(steps), 0]
    for l in macro:
        lo, hi = topo.cell_off[l], topo.cell_off[l + 1]
        a, b, c, e = (last[k, lo:hi].to(st) for k in range(4))
        lanes[l]._assign("curr", a, b, c, e)
        lanes[l]._assign("next", a, b, c, e)
    # ---- unpack: vehicles (one host copy of the small bookkeeping columns)
    book = auxh[:, 0, topo.A_FRONT:topo.A_CAP].detach().round().long().cpu().numpy()      # [T+1][front | count | nsp][ML]
    front, count, nsp = book[:, :ML], book[:, ML:2 * ML], book[:, 2 * ML:3 * ML]
    events = []                                   # (step, spawning macro lane, micro lane) in the reference's order
    for t in range(int(steps)):
        for m in np.nonzero(nsp[t + 1] > nsp[t])[0]:
            l = topo.micro[m]
            srcs = [a for a, b in cap_pairs if b == l and mroute.get_next_lane(a) == l]
            for a in sorted(srcs)[:int(nsp[t + 1][m] - nsp[t][m])]:
                events.append((t, a, l))
    events.sort(key=lambda e: (e[0], e[1]))
    spawned = {l: [] for l in topo.micro}
    for t, a, l in events:
        nv = MicroVehicle.default_micro_vehicle(lanes[l].speed_limit)
        nv.id = net._num_vehicle
        net._num_vehicle += 1
        net._vehicle[nv.id] = nv
        net._micro_route[nv.id] = net.create_random_route(l)          # replays the reference's np.random draws
        spawned[l].append(nv)
    A = auxh[int(steps), 0]
    for m, l in enumerate(topo.micro):
        n, f = int(count[-1][m]), int(front[-1][m])
        seq = olds[m] + spawned[l]
        gone = len(seq) - n
        assert gone >= 0, "vehicle bookkeeping of the fused hybrid rollout is inconsistent"
        keep = seq[gone:]                                             # head first
        idx = torch.tensor([(f + k) % cap for k in range(n)], dtype=torch.long, device=dev)
        base = m * cap
        pv = A[topo.A_P + base:topo.A_P + base + cap][idx].to(st)
        vv = A[topo.A_V + base:topo.A_V + base + cap][idx].to(st)
        av = A[topo.A_A + base:topo.A_A + base + cap][idx].to(st)
        lane = lanes[l]
        lane._curr_vehicle = list(reversed(keep))
        if n:
            for mv, a_ in zip(keep, av.unbind(0)):
                mv.a = a_
            lane._hand_out(torch.flip(pv, dims=[0]), torch.flip(vv, dims=[0]))
            lane._set_next(lane._curr_views[0].vec, lane._curr_views[1].vec)
        else:
            lane._curr_views = None
            lane._set_next(torch.zeros(0, dtype=st, device=dev), torch.zeros(0, dtype=st, device=dev))
        hd = headh[int(steps) - 1, 0, m]
        lane.head_position_delta, lane.head_speed_delta = hd[0].to(st), hd[1].to(st)
    for j, (a, b) in enumerate(cap_pairs):
        if mroute.get_next_lane(a) == b:
            lanes[a]._flux_capacitor[b] = A[topo.A_CAP + j].to(st)
    rt.check_flags()


def _walk(topo, l):
    path = [l]
    while topo.kind[path[-1]] and topo.next[path[-1]]:
        path.append(topo.next[path[-1]][0])
    return path
