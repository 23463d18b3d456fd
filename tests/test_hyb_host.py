"""Host side of the hybrid-network rollout (no GPU): lane graph tables, vehicle-route enumeration, conversion groups,
aux-row layout, and that the product path refuses to run without CUDA."""
import numpy as np
import pytest
import torch

from hyb_cases import build, fixture_case, spawn_routes


@pytest.mark.parametrize("tag", ["h", "g"])
def test_topology_tables_cover_reference_routes(tag):
    G = fixture_case(tag)
    grid, topo = build(G, "cpu")
    assert topo.L == len(G["kind"]) == 144 and topo.ML == 16 and topo.NC == G["r0"].shape[0]
    # every route the live reference drew for a spawned vehicle is in the enumerated table (cut after its first macro lane)
    nveh = int(G["vid"].max()) + 1
    for v in range(nveh):
        path = [int(x) for x in G["vroute"][v] if x >= 0]
        rid = topo.route_id(path)
        cut = topo.routes[rid]
        assert list(cut) == path[:len(cut)] and (not topo.kind[cut[-1]] or len(cut) == len(path))
        assert all(topo.kind[l] for l in cut[:-1])
    tab = spawn_routes(G, topo)
    assert tab.shape[0] == topo.ML and tab.max() < len(topo.routes)
    # conversion groups partition the lanes that take part in conversions, in lane-id order inside a group
    flat = [l for g in topo.groups for l in g]
    assert len(flat) == len(set(flat)) and all(g == sorted(g) for g in topo.groups)
    assert set(topo.micro) <= set(flat)
    for l in range(topo.L):
        if not topo.kind[l] and any(topo.kind[x] for x in topo.next[l]):
            assert l in flat
    # capacitors: one per (macro lane, next micro lane) pair (_macro_lane.py:215-225)
    assert topo.NCAP == sum(1 for l in range(topo.L) if not topo.kind[l] for x in topo.next[l] if topo.kind[x])
    assert topo.MAXT == int(np.ceil(topo.veh_len / min(d for l, d in enumerate(topo.cell_length) if not topo.kind[l]))) + 1


def test_aux_layout_and_initial_rows():
    G = fixture_case("h")
    _, topo = build(G, "cpu", veh_cap=4)
    s = topo.ML * topo.veh_cap
    assert topo.AUX == 6 * s + 3 * topo.ML + topo.NCAP + 3
    p0 = torch.arange(2 * s, dtype=torch.float64).reshape(2, topo.ML, topo.veh_cap)
    cnt = [1] * topo.ML
    aux = topo.make_aux0(2, torch.float64, p0=p0, count0=cnt)
    assert aux.shape == (2, topo.AUX)
    assert torch.equal(aux[:, topo.A_P:topo.A_V], p0.reshape(2, -1))
    assert torch.equal(aux[0, topo.A_CNT:topo.A_CNT + topo.ML], torch.ones(topo.ML, dtype=torch.float64))
    assert float(aux[:, topo.A_CAP:].abs().sum()) == 0.0
    with pytest.raises(AssertionError):
        topo.make_aux0(1, torch.float64, count0=[topo.veh_cap + 1] * topo.ML)


def test_random_spawn_routes_follow_the_graph():
    G = fixture_case("h")
    _, topo = build(G, "cpu")
    gen = torch.Generator().manual_seed(7)
    tab = topo.random_spawn_routes(2, 3, gen)
    assert tab.shape == (2, topo.ML, 3)
    for b in range(2):
        for m, l0 in enumerate(topo.micro):
            for k in range(3):
                path = topo.routes[int(tab[b, m, k])]
                assert path[0] == l0
                assert all(path[i + 1] in topo.next[path[i]] for i in range(len(path) - 1))


def test_no_cpu_fallback():
    from dhts_b200.hybrid_network import hybrid_rollout
    G = fixture_case("h")
    _, topo = build(G, "cpu")
    r0 = torch.tensor(G["r0"][None]); u0 = torch.tensor(G["u0"][None])
    with pytest.raises(Exception):
        hybrid_rollout(topo, r0, u0, 60.0, 1 / 30, 2, sig=torch.zeros(1, 2, topo.L, dtype=torch.float64),
                       incoming=torch.zeros(1, 2, topo.L, dtype=torch.float64),
                       route=torch.tensor(G["route"][:2], dtype=torch.int32), spawn_route=torch.zeros(topo.ML, 1, dtype=torch.int32))
