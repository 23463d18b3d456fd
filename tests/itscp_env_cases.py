"""Shared pieces of the headless-ITSCP tests: the full config-4 episode frozen from the live reference
(oracle/gen_golden_c4.py) and an env configured as run_itscp_hybrid.sh configures the reference's."""
import numpy as np
import torch

from conftest import golden


def c4_fixture():
    g = golden("itscp_c4_fp64")
    return {k[2:]: g[k] for k in g.files}


def c4_env(G, device, mode="hybrid"):
    """example/control/itscp/run.py:48-59 with run_itscp_hybrid.sh's arguments; schedule and MacroRoute draws replaced by
    the fixture's (they are random inputs)."""
    from dhts_b200.itscp_env import ItscpEnv
    inc = G["incoming"]
    env = ItscpEnv(device=device)
    env.schedule_callback = lambda ids, T: {i: inc[:T, k].tolist() for k, i in enumerate(ids)}
    env.config.update(num_intersection=3, lane_length=5.0, num_lane=1, render=False, policy_length=20, signal_length=4,
                      mode=mode, speed_limit=60.0, random_seed=5)
    env.reset()
    env.macro_route_schedule = torch.tensor(G["route"], dtype=torch.int32, device=device)
    return env


def c4_spawn_routes(G, topo, max_spawn=16):
    tab = np.zeros((topo.ML, max_spawn), dtype=np.int32)
    nsp = [0] * topo.ML
    for row in G["vroute"]:
        path = [int(x) for x in row if x >= 0]
        if not path:
            continue
        m = topo.mic_of[path[0]]
        tab[m, nsp[m]] = topo.route_id(path)
        nsp[m] += 1
    return tab
