"""TEST INFRASTRUCTURE: freeze a rollout of the LIVE reference (fp64 through the no-edit dtype rebinding of SURVEY
App. C) in which the lane is coupled to the outside at EVERY step -- new ghost cells before each step (what
RoadNetwork.setup_macro_boundary does for a lane inside a network, road_network.py:364-387) and a loss that reads
the state before each step (like the ITSCP queue length, example/control/itscp/_env.py:662-742) -- into
tests/golden/arz_perstep_fp64.npz.  Run in THIS container (needs /root/reference):

    python oracle/gen_golden_perstep.py
"""
import os
import sys

import numpy as np
import torch as th

REF = os.environ.get("DHTS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import switch_fp64  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "arz_perstep_fp64.npz")


def main():
    switch_fp64()
    from road.lane.dmacro_lane import dMacroLane
    from road.network.road_network import RoadNetwork
    rng = np.random.default_rng(424242)
    B, N, T, dx, umax, dt = 3, 24, 60, 5.0, 30.0, 0.01
    r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 1, (B, N)) * umax
    gh = np.stack([rng.uniform(0, 1, (T, B, 2)), rng.uniform(0, 1, (T, B, 2)) * umax], -1)       # [T,B,2,(r,u)]
    w_r = rng.normal(size=(T, B, N)); w_u = rng.normal(size=(T, B, N)) / umax                   # per-step loss weights
    wT_r = rng.normal(size=(B, N)); wT_u = rng.normal(size=(B, N)) / umax
    out = dict(rT=[], uT=[], g_r0=[], g_u0=[], g_ghost=[], loss=[], r_hist=[], u_hist=[])
    for b in range(B):
        tr = th.tensor(r0[b], dtype=th.float64, requires_grad=True)
        tu = th.tensor(u0[b], dtype=th.float64, requires_grad=True)
        tg = th.tensor(gh[:, b], dtype=th.float64, requires_grad=True)           # [T,2,2]
        lane = dMacroLane(0, N * dx, umax, dx)
        lane.set_state_vector_u(tr, tu)
        net = RoadNetwork(umax); net.add_lane(lane)
        loss = 0.0
        rh, uh = [], []
        for t in range(T):
            r, y, u = lane.get_state_vector()
            rh.append(r.detach().numpy().copy()); uh.append(u.detach().numpy().copy())
            loss = loss + (r * th.tensor(w_r[t, b])).sum() + (u * th.tensor(w_u[t, b])).sum()
            lane.set_leftmost_cell(tg[t, 0, 0], tg[t, 0, 1]); lane.set_rightmost_cell(tg[t, 1, 0], tg[t, 1, 1])
            net.forward(dt, True)
        r, y, u = lane.get_state_vector()
        loss = loss + (r * th.tensor(wT_r[b])).sum() + (u * th.tensor(wT_u[b])).sum()
        loss.backward()
        out["rT"].append(r.detach().numpy()); out["uT"].append(u.detach().numpy())
        out["g_r0"].append(tr.grad.numpy()); out["g_u0"].append(tu.grad.numpy()); out["g_ghost"].append(tg.grad.numpy())
        out["loss"].append(float(loss)); out["r_hist"].append(np.stack(rh)); out["u_hist"].append(np.stack(uh))
    o = {k: np.stack(v).astype(np.float64) for k, v in out.items()}
    o["g_ghost"] = np.transpose(o["g_ghost"], (1, 0, 2, 3))                      # [T,B,2,2]
    o["r_hist"] = np.transpose(o["r_hist"], (1, 0, 2)); o["u_hist"] = np.transpose(o["u_hist"], (1, 0, 2))
    np.savez(OUT, r0=r0, u0=u0, ghost_ru=gh, w_r=w_r, w_u=w_u, wT_r=wT_r, wT_u=wT_u, B=B, N=N, T=T, dx=dx, umax=umax, dt=dt, **o)
    print("wrote", OUT, "loss", o["loss"])


if __name__ == "__main__":
    main()
