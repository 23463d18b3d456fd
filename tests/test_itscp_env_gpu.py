"""GPU: the headless ITSCP env (dhts_b200.itscp_env) end to end against ONE FULL config-4 episode of the live reference
(oracle/gen_golden_c4.py: run_itscp_hybrid.sh's configuration, 600 frames from empty lanes, problem_1 inflow) -- the
fused hybrid rollout through the C ABI, the env's own running-mean sigmoid constants and queue reward, and the gradient
of the reward wrt the 45 actions.  fp64; tolerances written at each assert (north-star bar: rtol 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import relerr
from itscp_env_cases import c4_env, c4_fixture, c4_spawn_routes

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _episode(G, dev):
    env = c4_env(G, dev)
    topo = env.topo
    sp = torch.tensor(c4_spawn_routes(G, topo), dtype=torch.int32, device=dev)
    action = torch.tensor(G["action"], dtype=F64, device=dev, requires_grad=True)
    reward = env.rollout(action[None], True, spawn_routes=sp, keep_states=True)[0]
    return env, action, reward, env.last["states"]


def test_config4_episode_matches_live_reference(dev):
    G = c4_fixture()
    T, every = int(G["T"]), int(G["every"])
    env, action, reward, st = _episode(G, dev)
    bits, _ = env.flags.check(quiet_collisions=True)
    # discrete events: vehicles per micro lane at every frame (12 spawns, hand-offs, absorptions)
    assert (st.count[:, 0].cpu().numpy() == G["vcnt"]).all() and G["vcnt"].max() >= 1
    cells = st.cells[::every, 0].detach().cpu().numpy()
    assert np.abs(cells - G["hist_thin"]).max() < 1e-8
    p, v, a, valid = st.by_rank()
    ours = torch.stack([p, v, a], -1)[::every, 0].detach().cpu().numpy()
    mask = np.arange(ours.shape[2])[None, None] < G["vcnt"][::every][..., None]
    assert np.abs((ours - G["veh_thin"][:, :, :ours.shape[2]])[mask]).max() < 1e-8
    # reward with the env's OWN running-mean constants (the fixture stores none)
    assert abs(float(reward) - float(G["reward"])) < 1e-8 * abs(float(G["reward"]))
    t = lambda x: torch.tensor(x, dtype=F64, device=dev)
    wv = t(G["w_veh"])[:, :p.shape[-1]]
    term = (st.cells[T, 0, 0] * t(G["w_r"])).sum() + (st.cells[T, 0, 2] * t(G["w_u"])).sum() + \
        ((p[T, 0] * wv[..., 0] + v[T, 0] * wv[..., 1] + a[T, 0] * wv[..., 2]) * valid[T, 0].to(F64)).sum()
    assert abs(float(term) - float(G["term"])) < 1e-8 * abs(float(G["term"]))
    (reward + term).backward()
    env.flags.check(quiet_collisions=True)
    assert relerr(action.grad.cpu().numpy(), G["g_action"]) < 1e-6


def test_step_api_and_hard_evaluation(dev):
    """env.step (trainer.py:188): 0-dim differentiable reward; differentiable=False takes hard signals and the hard
    queue test (_env.py:576-586, 925-962) and carries no gradient."""
    G = c4_fixture()
    env = c4_env(G, dev)
    sp = torch.tensor(c4_spawn_routes(G, env.topo), dtype=torch.int32, device=dev)
    env._spawn_routes = sp[None]
    a = torch.tensor(G["action"], dtype=torch.float32, requires_grad=True)      # CPU fp32 action, as the controller emits
    obs, reward, terminal, info = env.step(a, True)
    assert terminal and obs.shape == (1440,) and reward.dim() == 0 and len(info["img"]) == 600
    assert abs(float(reward) - float(G["reward"])) < 1e-5 * abs(float(G["reward"]))     # fp32 action rounding
    reward.backward()
    assert a.grad is not None and a.grad.shape == (45,) and bool(torch.isfinite(a.grad).all()) and float(a.grad.abs().max()) > 0
    env2 = c4_env(G, dev); env2._spawn_routes = sp[None]
    _, r_hard, _, _ = env2.step(a.detach(), False)
    assert not r_hard.requires_grad and float(r_hard) <= 0 and np.isfinite(float(r_hard))


def test_macro_mode_matches_itscp_batch(dev):
    """mode='macro': the env's reward equals ItscpBatch's exact-constant reward on the same inputs, replica by replica."""
    from dhts_b200.itscp import ItscpBatch
    G = c4_fixture()
    env = c4_env(G, dev, mode="macro")
    assert not env.hybrid and env.topo.NC > 256
    rng = np.random.default_rng(3)
    act = torch.tensor(rng.uniform(0.2, 0.8, (3, 45)), dtype=F64, device=dev, requires_grad=True)
    r = env.rollout(act, True)
    batch = ItscpBatch(env.grid, dev, speed_limit=60.0, signal_length=4.0)
    act2 = act.detach().clone().requires_grad_()
    r2, _ = batch.rollout(act2, env.incoming()[None].expand(3, -1, -1).contiguous(), env.macro_route_schedule, 600)
    assert relerr(r.detach().cpu().numpy(), r2.detach().cpu().numpy()) < 1e-12
    r.sum().backward(); r2.sum().backward()
    assert relerr(act.grad.cpu().numpy(), act2.grad.cpu().numpy()) < 1e-10


def test_trainer_epoch_on_the_fused_path(dev, tmp_path):
    """Two epochs of the reference's training loop (run.py:65-70 shape: eval, train, save) with 3 batched episodes per
    epoch: finite controller gradients through controller -> action -> signals -> hybrid rollout -> queue reward."""
    from dhts_b200.control import Trainer
    G = c4_fixture()
    env = c4_env(G, dev)
    env.config["policy_length"] = 4; env.config["signal_length"] = 2           # 120 frames, 18 actions
    np.random.seed(5); env.reset()
    torch.manual_seed(0)
    tr = Trainer(env, lr=1e-4, tensorboard=False)
    before = [p.detach().clone() for p in tr.controller.parameters()]
    losses = tr.train(3, 2, 1, 1, str(tmp_path / "trial_0"))
    assert len(losses) == 2 and all(np.isfinite(losses)) and losses[0] > 0
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in tr.controller.parameters())
    assert any(float((p - q).abs().max()) > 0 for p, q in zip(tr.controller.parameters(), before))
    assert len(open(str(tmp_path / "trial_0" / "eval.txt")).read().splitlines()) == 2


@pytest.mark.parametrize("n,num_lane,lane_length,mode", [(4, 1, 20.0, "hybrid"), (2, 2, 10.0, "macro"), (3, 2, 5.0, "hybrid")])
def test_other_grids_run_clean(dev, n, num_lane, lane_length, mode):
    """Grids the fixtures do not cover (4x4 with a 2x2 micro core where vehicles hand over micro -> micro; two lanes per
    road; macro mode): no CFL / route / overflow flags, finite non-zero action gradients, hard evaluation finite."""
    from dhts_b200.itscp_env import ItscpEnv, problem_2
    np.random.seed(9)
    env = ItscpEnv(device=dev)
    env.schedule_callback = problem_2
    env.config.update(num_intersection=n, num_lane=num_lane, lane_length=lane_length, policy_length=6, signal_length=2,
                      mode=mode, speed_limit=60.0, max_spawn=32)
    env.reset()
    assert env.action_size() == 3 * n * n and env.hybrid == (mode == "hybrid" and n >= 3)
    g = torch.Generator().manual_seed(1)
    R = 3
    act = (0.3 + 0.4 * torch.rand((R, env.action_size()), generator=g)).to(dev).double().requires_grad_()
    env.resample_spawn_routes(R, g)
    r = env.rollout(act, True, keep_states=True)
    r.sum().backward()
    bits, ncol = env.flags.check(quiet_collisions=True)
    assert (bits & ~4) == 0
    assert bool(torch.isfinite(r).all()) and float(r.max()) < 0
    assert bool(torch.isfinite(act.grad).all()) and float(act.grad.abs().max()) > 0
    if env.hybrid:
        st = env.last["states"]
        assert int(st.count.max()) >= 1, "vehicles must have been spawned"
    with torch.no_grad():
        rh = env.rollout(act.detach(), False)
    assert bool(torch.isfinite(rh).all())


def test_run_itscp_cli_writes_the_reference_tree(dev, tmp_path):
    """python -m dhts_b200.run_itscp with run.py's arguments: <out>/trial_<k>/{eval.txt, model.zip, best/model.zip}."""
    import os
    from dhts_b200.run_itscp import main
    out = str(tmp_path / "hybrid_0")
    curves = main(["--mode=hybrid", "--problem=2", "--n_trial=2", "--n_intersection=3", "--n_lane=1", "--lane_length=5",
                   "--speed_limit=60", "--simulation_length=4", "--signal_length=2", "--n_episode=3", "--lr=1e-4",
                   "--episodes_per_epoch=2", "--seed=3", "--out", out])
    assert len(curves) == 2 and all(len(c) == 4 and np.isfinite(c).all() for c in curves)       # n_episode + 1 epochs (run.py:70)
    for k in range(2):
        d = os.path.join(out, "trial_%d" % k)
        assert os.path.exists(d + "/model.zip") and os.path.exists(d + "/best/model.zip")
        assert len(open(d + "/eval.txt").read().split()) == 4                                   # evaluation every max(3 // 10, 1) = 1 epoch
