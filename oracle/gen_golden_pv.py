"""TEST INFRASTRUCTURE: freeze a hybrid chain with HETEROGENEOUS vehicles from the LIVE reference into tests/golden/.

Run in THIS container only (needs /root/reference):

    python oracle/gen_golden_pv.py

macro(10) -> micro -> micro -> macro(10) through ``RoadNetwork.forward`` + ``Conversion.*`` (road_network.py:79-173,
conversion.py:15-215), dMacroLane / dMicroLane operators.  The two micro lanes start with vehicles made by
``MicroVehicle.random_micro_vehicle`` (road/vehicle/micro_vehicle.py:74-122): every vehicle has its own a_max, a_pref,
target speed, minimum gap and time headway, which it keeps when it is handed from micro lane 1 to micro lane 2
(conversion.py:174-200); vehicles spawned from the macro lane are ``default_micro_vehicle``s (conversion.py:53-57).
Loss on the last lane's cells and on the vehicles still on the micro lanes; gradients wrt the cells of lane 0 and wrt the
initial vehicle positions / speeds.

Fixture: tests/golden/hybrid_chain_pv_fp64.npz.
"""
import os
import sys

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DHTS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(1, HERE)

from gen_golden import switch_fp64, f  # noqa: E402


def main():
    switch_fp64()
    from road.lane.dmacro_lane import dMacroLane
    from road.lane.dmicro_lane import dMicroLane
    from road.network.road_network import RoadNetwork
    from road.network.route import MicroRoute
    from road.vehicle.micro_vehicle import MicroVehicle
    dtype = th.float64
    rng = np.random.default_rng(77)
    T = int(os.environ.get("PV_T", "600"))
    N, dx, umax, dt = 10, 5.0, 30.0, 0.01
    Llen = N * dx
    r0 = rng.uniform(0.3, 1.0, (2, N)); u0 = rng.uniform(0.3, 1.0, (2, N)) * umax
    gh = np.stack([rng.uniform(0.3, 1, 4), rng.uniform(0.3, 1, 4) * umax], -1)   # [4,(r,u)]
    np.random.seed(4321)
    net = RoadNetwork(umax)
    t0r = th.tensor(r0[0], dtype=dtype, requires_grad=True); t0u = th.tensor(u0[0], dtype=dtype, requires_grad=True)
    t3r = th.tensor(r0[1], dtype=dtype, requires_grad=True); t3u = th.tensor(u0[1], dtype=dtype, requires_grad=True)
    l0 = dMacroLane(0, Llen, umax, dx); l0.set_state_vector_u(t0r, t0u)
    l0.set_leftmost_cell(th.tensor(gh[0, 0], dtype=dtype), th.tensor(gh[0, 1], dtype=dtype))
    l0.set_rightmost_cell(th.tensor(gh[1, 0], dtype=dtype), th.tensor(gh[1, 1], dtype=dtype))
    net.add_lane(l0)
    l1 = dMicroLane(1, Llen, umax); net.add_lane(l1)
    l2 = dMicroLane(2, Llen, umax); net.add_lane(l2)
    l3 = dMacroLane(3, Llen, umax, dx); l3.set_state_vector_u(t3r, t3u)
    l3.set_leftmost_cell(th.tensor(gh[2, 0], dtype=dtype), th.tensor(gh[2, 1], dtype=dtype))
    l3.set_rightmost_cell(th.tensor(gh[3, 0], dtype=dtype), th.tensor(gh[3, 1], dtype=dtype))
    net.add_lane(l3)
    net.connect_lane(0, 1); net.connect_lane(1, 2); net.connect_lane(2, 3)
    # initial vehicles: 3 on lane 1, 2 on lane 2, head first in the arrays below
    pos = {1: [38.0, 24.0, 9.0], 2: [30.0, 12.0]}
    spd = {1: [11.0, 14.0, 9.0], 2: [16.0, 8.0]}
    par, p_t, v_t = {}, {}, {}
    for lane_id in (1, 2):
        par[lane_id] = []; p_t[lane_id] = []; v_t[lane_id] = []
        for p, v in zip(pos[lane_id], spd[lane_id]):
            mv = MicroVehicle.random_micro_vehicle(umax)
            pt = th.tensor(p, dtype=dtype, requires_grad=True); vt = th.tensor(v, dtype=dtype, requires_grad=True)
            mv.position = pt; mv.speed = vt
            net.add_vehicle(mv, MicroRoute([lane_id] + list(range(lane_id + 1, 4))))
            par[lane_id].append([mv.accel_max, mv.accel_pref, mv.target_speed, mv.min_space, mv.time_pref, mv.length])
            p_t[lane_id].append(pt); v_t[lane_id].append(vt)
    net.macro_route = net.create_random_macro_route()
    cnt_hist = []; nspawn = []
    for t in range(T):
        net.forward(dt, True)
        cnt_hist.append([l1.num_vehicle(), l2.num_vehicle()]); nspawn.append(net.num_vehicle)
    s0 = l0.get_state_vector(); s3 = l3.get_state_vector()
    w_r = rng.normal(size=N); w_u = rng.normal(size=N) / umax
    loss = (s3[0] * th.tensor(w_r, dtype=dtype)).sum() + (s3[2] * th.tensor(w_u, dtype=dtype)).sum()
    w_veh = rng.normal(size=8)
    final = {}
    for lane in (l1, l2):
        rows = []
        for i, mv in enumerate(lane.curr_vehicle):      # tail first
            loss = loss + w_veh[2 * i % 8] * mv.position * 0.01 + w_veh[(2 * i + 1) % 8] * mv.speed * 0.01
            rows.append([f(mv.position), f(mv.speed), f(mv.a), mv.accel_max, mv.time_pref])
        final[lane.id] = np.array(rows).reshape(-1, 5)
    loss.backward()
    z = lambda t: np.zeros(N) if t.grad is None else t.grad.numpy().astype(np.float64)
    zs = lambda ts: np.array([0.0 if t.grad is None else f(t.grad) for t in ts])
    out = dict(r0=r0, u0=u0, ghost_ru=gh, w_r=w_r, w_u=w_u, w_veh=w_veh, N=N, dx=dx, umax=umax, dt=dt, T=T,
               pos1=np.array(pos[1]), spd1=np.array(spd[1]), par1=np.array(par[1]), pos2=np.array(pos[2]), spd2=np.array(spd[2]),
               par2=np.array(par[2]), cnt_hist=np.array(cnt_hist), nspawn_hist=np.array(nspawn),
               lane0=np.stack([x.detach().numpy().astype(np.float64) for x in s0]),
               lane3=np.stack([x.detach().numpy().astype(np.float64) for x in s3]),
               veh1=final[1], veh2=final[2], loss=f(loss), g_r0_lane0=z(t0r), g_u0_lane0=z(t0u), g_r0_lane3=z(t3r), g_u0_lane3=z(t3u),
               g_pos1=zs(p_t[1]), g_spd1=zs(v_t[1]), g_pos2=zs(p_t[2]), g_spd2=zs(v_t[2]))
    print("chain_pv: vehicles", net.num_vehicle, "on lanes", l1.num_vehicle(), l2.num_vehicle(), "loss", out["loss"],
          "|g lane0|", np.abs(out["g_r0_lane0"]).max(), "|g pos|", np.abs(out["g_pos1"]).max(), np.abs(out["g_pos2"]).max())
    path = os.path.join(OUT, "hybrid_chain_pv_fp64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
