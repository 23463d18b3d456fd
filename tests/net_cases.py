"""Shared pieces of the connected-network tests (CPU: checker vs live-reference fixtures; GPU: kernels vs both)."""
import numpy as np

from conftest import golden


def fixture_case(tag):
    g = golden("itscp_macro_fp64")
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}


def grid_of(G):
    from dhts_b200.itscp import ItscpGrid
    return ItscpGrid(int(G["num_intersection"]), int(G["num_lane"]), float(G["lane_length"]), float(G["cell_length"]))


def reward_and_injection(hist, off, dx, k, dt, veh_len, static_speed):
    """Queue reward (example/control/itscp/_env.py:618-648,770-797) of stored states with given per-sample sigmoid
    constants k [T, NC], and its gradient wrt the states after each step, in numpy."""
    T = hist.shape[0] - 1
    NC = hist.shape[2]
    g = np.zeros((T, 3, NC)); rew = 0.0
    for t in range(T):
        r, u = hist[t + 1, 0], hist[t + 1, 2]
        for l in range(len(dx)):
            s = slice(off[l], off[l + 1])
            w = dx[l] / veh_len
            zr = (static_speed - u[s]) * k[t, s]
            z = np.clip(zr, -16, 16); sg = 1 / (1 + np.exp(-z))
            q = (sg * r[s] * w).sum()
            rew += -q * q * dt
            g[t, 0, s] += -2 * q * dt * sg * w
            g[t, 2, s] += -2 * q * dt * r[s] * w * sg * (1 - sg) * (-k[t, s]) * ((zr >= -16) & (zr <= 16))
    return rew, g


def random_network(rng, L, max_cells=4, extra_links=0.5):
    """A random connected-ish lane graph: a chain backbone plus random forward links, so that lanes have 0, 1 or
    several neighbours per side.  Returns (num_cell, dx, links)."""
    num_cell = rng.integers(1, max_cells + 1, L).tolist()
    dx = rng.uniform(4.0, 6.0, L).tolist()
    links = set()
    order = rng.permutation(L)
    for a, b in zip(order[:-1], order[1:]):
        if rng.uniform() < 0.8:
            links.add((int(a), int(b)))
    for _ in range(int(extra_links * L)):
        i, j = sorted(rng.choice(L, 2, replace=False).tolist())
        links.add((int(order[i]), int(order[j])))
    return num_cell, dx, sorted(links)


def random_routes(rng, prev, nxt, T):
    """Per step and lane side a selected neighbour: always one when the lane has several (the reference raises
    KeyError otherwise), the only neighbour or none (-1: ITSCP treats it as a red light) when it has one."""
    L = len(prev)
    tab = -np.ones((T, 2, L), dtype=np.int32)
    for t in range(T):
        for side, lists in enumerate((prev, nxt)):
            for l in range(L):
                if len(lists[l]) > 1:
                    tab[t, side, l] = rng.choice(lists[l])
                elif len(lists[l]) == 1 and rng.uniform() < 0.7:
                    tab[t, side, l] = lists[l][0]
    return tab
