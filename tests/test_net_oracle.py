"""CPU: the network checker (oracle/net_oracle.py) and the ITSCP host logic against fixtures frozen from the LIVE
reference's ItscpRoadNetwork in macro mode (oracle/gen_golden_net.py)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from net_cases import fixture_case, grid_of, reward_and_injection
from oracle import net_oracle as NO


@pytest.mark.parametrize("tag", ["a", "b"])
def test_checker_matches_live_reference(tag):
    G = fixture_case(tag)
    grid = grid_of(G)
    net = NO.Net(grid.num_cell, grid.dx, grid.links, 1)
    T, umax, dt = int(G["T"]), float(G["umax"]), float(G["dt"])
    rew, gst = reward_and_injection(G["hist"], net.off, net.dx, G["kconst"], dt, float(G["veh_len"]), float(G["static_speed"]))
    assert abs(rew - float(G["reward"])) < 1e-12 * max(1.0, abs(rew))
    gst[T - 1, 0] += G["w_r"]; gst[T - 1, 2] += G["w_u"]
    o = NO.rollout(net, G["r0"], G["u0"], umax, dt, T, sig=G["sig"], incoming=G["incoming"], route=G["route"], soft=True,
                   g_states=gst, want_grad=True)
    assert o["cfl"] == 0
    assert np.abs(o["hist"] - G["hist"]).max() < 1e-12
    tensor_sig = np.array([(i.loc != "mid" and i.approaching) for i in grid.lanes])      # the others are the constant 1.0
    assert relerr(o["g_sig"][:, tensor_sig], G["g_sig"][:, tensor_sig]) < 1e-9
    for k in ("g_inc", "g_r0", "g_u0"):
        assert relerr(o[k], G[k]) < 1e-7, k


@pytest.mark.parametrize("tag", ["a", "b"])
def test_itscp_host_logic_matches_live_reference(tag):
    """Lane graph discretisation, per-frame lane signals and the running-mean sigmoid constants."""
    from dhts_b200.itscp import queue_constants
    G = fixture_case(tag)
    grid = grid_of(G)
    T = int(G["T"])
    assert sum(grid.num_cell) == G["hist"].shape[2]
    a = torch.tensor(G["action"]).unsqueeze(0)
    sig = grid.signals(a, T, int(G["frames_per_signal"]), soft=True)[0].numpy()
    assert np.abs(sig - G["sig"]).max() < 1e-14
    k = queue_constants(torch.tensor(G["hist"][1:, 2]), float(G["static_speed"])).numpy()
    assert relerr(k, G["kconst"]) < 1e-6      # the reference rounds every sample to float32 (common/rms.py:12-13); we keep float64
    hard = grid.signals(a, T, int(G["frames_per_signal"]), soft=False)[0].numpy()
    assert set(np.unique(hard)) <= {0.0, 1.0} and np.all((hard > 0.5) == (G["sig"] > 0.5))


def test_running_mean_window():
    """The window of the running mean (common/rms.py) drops old samples one by one."""
    from dhts_b200.itscp import queue_constants
    rng = np.random.default_rng(3)
    u = rng.uniform(0, 30, (7, 13))
    k = queue_constants(torch.tensor(u), 0.2, window=20).numpy().ravel()
    d = (0.2 - u).ravel()
    ref = np.array([16.0 / abs(d[max(0, i - 19):i + 1].mean()) for i in range(d.size)])
    assert relerr(k, ref) < 1e-12
