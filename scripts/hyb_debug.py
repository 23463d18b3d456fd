"""Diagnostics: fused hybrid rollout vs the live-reference ITSCP hybrid fixture (first divergence, gradient errors)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import relerr  # noqa: E402
from hyb_cases import fixture_case, run_fixture  # noqa: E402

dev = torch.device("cuda:0")
for tag in sys.argv[1:] or ["h", "g"]:
    G = fixture_case(tag)
    o = run_fixture(G, dev)
    st, topo = o["st"], o["topo"]
    T = int(G["T"])
    print("== case", tag, "flags", o["flags"].read(), "groups", [len(g) for g in topo.groups], "routes", len(topo.routes))
    cells = st.cells[:, 0].detach().cpu().numpy()
    cnt = st.count[:, 0].cpu().numpy()
    p, v, a = (x[:, 0].detach().cpu().numpy() for x in (o["p"], o["v"], o["a"]))
    head = st.head[:, 0].detach().cpu().numpy()
    for t in range(T + 1):
        ec = np.abs(cells[t] - G["hist"][t]).max()
        bad_cnt = (cnt[t] != G["vcnt"][t]).any()
        ev = 0.0
        if not bad_cnt:
            for m in range(topo.ML):
                n = cnt[t, m]
                if n:
                    ev = max(ev, np.abs(np.stack([p[t, m, :n], v[t, m, :n], a[t, m, :n]], -1) - G["veh"][t, m, :n]).max())
        eh = np.abs(head[t] - G["head"][t]).max() if t < T else 0.0
        if ec > 1e-9 or ev > 1e-9 or bad_cnt or eh > 1e-7:
            print("first divergence at t =", t, "cells", ec, "veh", ev, "count mismatch", bad_cnt, "head", eh)
            if bad_cnt:
                print(" ours", cnt[t], "\n ref ", G["vcnt"][t])
            if eh > 1e-7:
                k = np.argmax(np.abs(head[t] - G["head"][t]).max(-1)); print(" head lane", topo.micro[k], head[t, k], G["head"][t, k])
            if ec > 1e-9:
                c = np.unravel_index(np.argmax(np.abs(cells[t] - G["hist"][t])), cells[t].shape); print(" cell", c, cells[t][c], G["hist"][t][c])
            break
    else:
        print("states match: cells", np.abs(cells - G["hist"]).max(), "spawned", int(G["vid"].max()) + 1)
    print("reward", float(o["reward"]), float(G["reward"]), "term", float(o["term"]), float(G["term"]))
    (o["reward"] + o["term"]).backward()
    print("flags after bwd", o["flags"].read())
    lanes = [l for l, info in enumerate(o["grid"].lanes) if info.loc != "mid" and info.approaching]
    print("g_action", relerr(o["action"].grad[0].cpu().numpy(), G["g_action"]))
    print("g_sig   ", relerr(o["sig"].grad[0].cpu().numpy()[:, lanes], G["g_sig"][:, lanes]))
    print("g_inc   ", relerr(o["inc"].grad[0].cpu().numpy(), G["g_inc"]))
    print("g_r0    ", relerr(o["r0"].grad[0].cpu().numpy(), G["g_r0"]))
    print("g_u0    ", relerr(o["u0"].grad[0].cpu().numpy(), G["g_u0"]))
