"""CPU: host logic of ITSCP MICRO mode (example/control/itscp/_env.py:143-219): a seeded ``reset``
draws the schedule, the per-frame MacroRoutes and the waiting routes of every lane with the reference's ``np.random``
calls in the reference's order, so it must reproduce what the live reference drew (frozen by
oracle/gen_golden_micro.py) and leave ``np.random`` in the same state."""
import numpy as np
import pytest

from itscp_micro_cases import CASES, micro_env, micro_fixture


@pytest.mark.parametrize("tag", ["m", "n"])
def test_reset_reproduces_the_reference_draws(tag):
    G = micro_fixture(tag)
    env = micro_env(tag, G, "cpu")
    assert env.micro_mode and env.hybrid and env.topo.NC == 0 and env.topo.ML == env.grid.L
    assert env.num_timestep == int(G["T"])
    # schedule of the boundary lanes (problem.py:5-68)
    inc = env.incoming().numpy()
    assert np.abs(inc - G["incoming"]).max() == 0.0
    # waiting routes of every lane, list order (_env.py:202-219, road_network.py:604-646)
    wait = G["wait"]
    L, K = wait.shape[:2]
    assert len(env.waiting_route) == L and all(len(r) == K for r in env.waiting_route)
    for l in range(L):
        for k in range(K):
            assert env.waiting_route[l][k] == [int(x) for x in wait[l, k] if x >= 0]
    # pop order: the k-th vehicle entering a lane rides the route at the END of what is left of the list
    ids = env._wait_ids.numpy()
    for l in range(L):
        assert [env.topo.routes[i] for i in ids[l]] == [tuple(r) for r in reversed(env.waiting_route[l])]
    # np.random is where the reference's reset left it
    st = np.random.get_state()
    assert int(st[2]) == int(G["rng_after_reset_pos"]) and (st[1] == G["rng_after_reset_key"]).all()
    # sources: exactly the lanes without predecessor
    assert [l for l, s in zip(env.topo.micro, env.topo.src) if s] == env.grid.boundary_lanes()
    assert env.veh_cap >= int(G["vcnt"].max()) + 1


def test_micro_mode_is_accepted_and_unknown_modes_are_not():
    from dhts_b200.itscp_env import ItscpEnv
    env = ItscpEnv(device="cpu")
    env.config["mode"] = "meso"
    with pytest.raises(ValueError):
        env.reset()
    assert set(CASES) == {"m", "n"}
