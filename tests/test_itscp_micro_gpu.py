"""GPU: ITSCP MICRO mode of the headless env (run_itscp_micro.sh) against episodes of the live
reference (oracle/gen_golden_micro.py): every lane a plain MicroLane, boundary lanes fed from the stochastic waiting
lists (_simulator.py:153-174), signal-blended head deltas, micro->micro hand-offs and drops, queue reward over the
vehicles with the running-mean sigmoid constant, gradient wrt the actions (the reference: plain autograd through
MicroLane.forward; here the adjoint kernel).  fp64; tolerances at each assert (north-star bar: rtol 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from itscp_micro_cases import micro_env, micro_fixture

pytestmark = pytest.mark.gpu
F64 = torch.float64


@pytest.mark.parametrize("tag", ["m", "n"])
def test_micro_episode_matches_live_reference(dev, tag):
    G = micro_fixture(tag)
    T = int(G["T"])
    env = micro_env(tag, G, dev)
    action = torch.tensor(G["action"], dtype=F64, device=dev, requires_grad=True)
    # the env takes its uniform draws from np.random, which the seeded reset left where the reference's reset did
    reward = env.rollout(action[None], True, keep_states=True)[0]
    bits, _ = env.flags.check(quiet_collisions=True)
    assert (bits & ~4) == 0
    st = env.last["states"]
    # discrete events: vehicles per lane at every frame (entries from the waiting lists, hand-offs, drops)
    assert (st.count[:, 0].cpu().numpy() == G["vcnt"]).all() and G["vcnt"].max() >= 4
    # the episode consumed exactly the reference's draws and np.random continues where the reference's does
    assert env.last_draws == len(G["draws"])
    assert float(np.random.random()) == float(G["rng_next"])
    p, v, a, valid = st.by_rank()
    ours = torch.stack([p, v, a], -1)[:, 0].detach().cpu().numpy()
    cap = min(ours.shape[2], G["veh"].shape[2])
    mask = np.arange(cap)[None, None] < G["vcnt"][..., None]
    assert np.abs((ours[:, :, :cap] - G["veh"][:, :, :cap])[mask]).max() < 1e-8
    # head deltas every lane with vehicles used (signal blend with the running-mean constant, _simulator.py:248-262)
    has = G["vcnt"][:-1] > 0      # vehicles before the step ... or one entered at the step
    hd = st.head[:, 0].detach().cpu().numpy()
    assert np.abs((hd - G["head"])[has]).max() < 1e-8 * max(1.0, np.abs(G["head"][has]).max())
    assert abs(float(reward) - float(G["reward"])) < 1e-8 * abs(float(G["reward"]))
    reward.backward()
    env.flags.check(quiet_collisions=True)
    assert relerr(action.grad.cpu().numpy(), G["g_action"]) < 1e-6


def test_explicit_draws_replicas_and_hard_evaluation(dev):
    """src_rand as an input: the recorded draws reproduce the episode whatever np.random holds; R replicas with the same
    draws are identical; differentiable=False takes hard signals and carries no gradient."""
    G = micro_fixture("m")
    env = micro_env("m", G, dev)
    np.random.seed(12345)
    draws = torch.tensor(G["draws"], dtype=F64, device=dev)
    act = torch.tensor(G["action"], dtype=F64, device=dev)
    R = 3
    r = env.rollout(act[None].expand(R, -1).contiguous(), True, src_rand=draws, keep_states=True)
    assert float((r - float(G["reward"])).abs().max()) < 1e-8 * abs(float(G["reward"]))
    assert bool((r == r[0]).all())
    st = env.last["states"]
    assert int(st.aux[-1, 0, env.topo.A_DRAW].round()) == len(G["draws"])
    # too few draws: the overflow flag, not a silent wrong answer
    with pytest.raises(RuntimeError):
        env.rollout(act[None], True, src_rand=draws[:50])
        env.flags.check(quiet_collisions=True)
    with torch.no_grad():
        rh = env.rollout(act[None], False, src_rand=torch.cat([draws, draws]))
    env.flags.check(quiet_collisions=True)
    assert not rh.requires_grad and np.isfinite(float(rh)) and float(rh) <= 0


def test_step_api_trainer_and_cli_in_micro_mode(dev, tmp_path):
    """env.step / the trainer loop / python -m dhts_b200.run_itscp --mode=micro (run_itscp_micro.sh's arguments, shortened)."""
    import os
    from dhts_b200.control import Trainer
    from dhts_b200.run_itscp import main
    G = micro_fixture("m")
    env = micro_env("m", G, dev)
    a = torch.tensor(G["action"], dtype=torch.float32, requires_grad=True)
    obs, reward, terminal, info = env.step(a, True)
    assert terminal and reward.dim() == 0 and len(info["img"]) == 300 and obs.shape == (10 * env.grid.L,)
    assert abs(float(reward) - float(G["reward"])) < 1e-4 * abs(float(G["reward"]))       # fp32 action rounding
    reward.backward()
    assert a.grad is not None and bool(torch.isfinite(a.grad).all()) and float(a.grad.abs().max()) > 0
    torch.manual_seed(0)
    tr = Trainer(env, lr=1e-4, tensorboard=False)
    losses = tr.train(1, 2, 1, 1, str(tmp_path / "trial_0"))
    assert len(losses) == 2 and all(np.isfinite(losses))
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in tr.controller.parameters())
    out = str(tmp_path / "micro_0")
    curves = main(["--mode=micro", "--problem=1", "--n_trial=1", "--n_intersection=1", "--n_lane=3", "--lane_length=30",
                   "--speed_limit=60", "--simulation_length=4", "--signal_length=2", "--n_episode=2", "--lr=1e-4", "--seed=3",
                   "--out", out])
    assert len(curves) == 1 and len(curves[0]) == 3 and np.isfinite(curves[0]).all()
    assert os.path.exists(out + "/trial_0/model.zip") and os.path.exists(out + "/trial_0/best/model.zip")
