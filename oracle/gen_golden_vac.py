"""TEST INFRASTRUCTURE: freeze a rollout of the LIVE reference over lanes with EMPTY and near-vacuum stretches.

Run in THIS container only (needs /root/reference):

    python oracle/gen_golden_vac.py        # dtype-proxied fp64 reference, writes tests/golden/arz_rollout_vac_fp64.npz

Lanes of 128 cells (so that the GPU kernels take the stored-outcome path with four cells per thread) whose initial
density has stretches of exact zeros and of values on both sides of eps = 1e-5, vacuum ghost cells on some lanes:
the vacuum branches of the case tree (_arz.py:225-322), the u_eq / flux_prime fix-ups below eps (darz.py:217-233)
and the clamps, through 60 steps of RoadNetwork.forward and the autograd chain back to the initial state and ghosts.
"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as G  # noqa: E402


def gen(dtype, rng, T):
    from road.lane.dmacro_lane import dMacroLane
    from road.network.road_network import RoadNetwork
    B, N, dx, umax, dt = 4, 128, 5.0, 30.0, 0.01
    r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 1, (B, N)) * umax
    for b in range(B):
        for _ in range(3):
            a = int(rng.integers(0, N - 8)); w = int(rng.integers(3, N // 6))
            r0[b, a:a + w] = 0.0
            a = int(rng.integers(0, N - 8)); w = int(rng.integers(3, N // 8))
            r0[b, a:a + w] = 10.0 ** rng.uniform(-7, -4, size=r0[b, a:a + w].shape)
    gh = np.stack([rng.uniform(0, 1, (B, 2)), rng.uniform(0, 1, (B, 2)) * umax], axis=-1)
    gh[0, 0, 0] = 0.0; gh[1, 1, 0] = 3e-6; gh[2, 0, 0] = 2e-6
    w_r = rng.normal(size=(B, N)); w_u = rng.normal(size=(B, N)) / umax
    res = dict(rT=[], yT=[], uT=[], g_r0=[], g_u0=[], g_ghost=[])
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
        for t in range(T):
            net.forward(dt, True)
        r, y, u = lane.get_state_vector()
        loss = (r * th.tensor(w_r[b], dtype=dtype)).sum() + (u * th.tensor(w_u[b], dtype=dtype)).sum()
        loss.backward()
        for k, v in (("rT", r), ("yT", y), ("uT", u)):
            res[k].append(v.detach().numpy().astype(np.float64))
        res["g_r0"].append(tr.grad.numpy().astype(np.float64)); res["g_u0"].append(tu.grad.numpy().astype(np.float64))
        res["g_ghost"].append(tg.grad.numpy().astype(np.float64))
    out = {k: np.stack(v) for k, v in res.items()}
    out.update(r0=r0, u0=u0, ghost_ru=gh, w_r=w_r, w_u=w_u, B=B, N=N, dx=dx, umax=umax, dt=dt, T=T)
    return out


if __name__ == "__main__":
    sys.path.insert(0, G.REF)
    sys.dont_write_bytecode = True
    G.switch_fp64()
    th.manual_seed(20221008); np.random.seed(20221008)
    data = gen(th.float64, np.random.default_rng(107), 60)
    assert (data["rT"] < 1e-5).any(), "no vacuum cell survived"
    np.savez_compressed(os.path.join(G.OUT, "arz_rollout_vac_fp64.npz"), **data)
    print("wrote arz_rollout_vac_fp64.npz: vacuum cells at the end:", int((data["rT"] < 1e-5).sum()),
          "max |g_r0|", float(np.abs(data["g_r0"]).max()))
