"""world_size-2 gloo test of the shard / reduce harness (dhts_b200.dist): the host-side logic of the N>1 path.
Each rank 'simulates' its lane shard with a deterministic per-lane function (the kernels need a GPU; what is tested
here is partitioning, CSR re-basing, the loss / shared-gradient reductions and lane-ordered gathering)."""
import os
import sys

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lane_fn(r, w):
    return (r * r).sum(dim=1) * w          # stands for a per-lane rollout loss with a shared parameter w


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=ws)
    from dhts_b200 import dist
    g = torch.Generator().manual_seed(3)
    B, N = 7, 5                                                   # uneven split on purpose: 4 + 3
    r = torch.rand((B, N), generator=g, dtype=torch.float64)
    sizes = torch.tensor([3, 0, 4, 1, 2, 6, 0, 5, 2])             # ragged micro lanes incl. empty ones
    off = torch.cat([torch.zeros(1, dtype=torch.int64), sizes.cumsum(0)]).to(torch.int32)
    V = int(off[-1])
    p = torch.arange(V, dtype=torch.float64); par = torch.arange(6 * V, dtype=torch.float64).reshape(6, V)
    head = torch.arange(2 * 9, dtype=torch.float64).reshape(9, 2)
    w = torch.tensor(0.5, dtype=torch.float64, requires_grad=True)
    (r_loc,) = dist.shard_lanes([r], rank, ws)
    off_loc, (p_loc, par_loc), (head_loc,) = dist.shard_csr(off, [p, par], [head], rank, ws)
    lane_loss = _lane_fn(r_loc, w)
    lane_loss.sum().backward()
    total = dist.reduce_losses(lane_loss.sum(), p_loc.sum())
    dist.reduce_shared_grads([w])
    gathered = dist.gather_lanes(lane_loss.detach(), B)
    lo, hi = dist.shard_range(9, rank, ws)
    q.put((rank, total.tolist(), float(w.grad), gathered.tolist(), off_loc.tolist(), p_loc.tolist(),
           par_loc.shape[1], head_loc.tolist(), (lo, hi)))
    td.destroy_process_group()


def test_two_rank_shards_reduce_to_the_unsharded_result():
    sys.path.insert(0, ROOT)
    from dhts_b200 import dist
    ws, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(ws))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # unsharded truth
    g = torch.Generator().manual_seed(3)
    r = torch.rand((7, 5), generator=g, dtype=torch.float64)
    w = torch.tensor(0.5, dtype=torch.float64, requires_grad=True)
    full = _lane_fn(r, w); full.sum().backward()
    sizes = [3, 0, 4, 1, 2, 6, 0, 5, 2]
    V = sum(sizes)
    for rank, total, wgrad, gathered, off_loc, p_loc, npar, head_loc, (lo, hi) in res:
        assert abs(total[0] - float(full.detach().sum())) < 1e-12 and total[1] == V * (V - 1) / 2
        assert abs(wgrad - float(w.grad)) < 1e-12
        assert gathered == full.detach().tolist()          # bitwise, in lane order
        assert off_loc[0] == 0 and off_loc[-1] == len(p_loc) == npar
        assert [b - a for a, b in zip(off_loc, off_loc[1:])] == sizes[lo:hi]
        assert head_loc == [[2.0 * l, 2.0 * l + 1] for l in range(lo, hi)]
    assert res[0][5] + res[1][5] == [float(i) for i in range(V)]          # vehicle slices tile the batch in order
    assert [dist.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert dist.world() == (0, 1)


def _trainer_worker(rank, ws, port, q, tmp):
    """Two ranks train one controller on 5 episodes per epoch (3 + 2) with a stub rollout whose reward depends on the
    episode's spawn-route draw; every rank must end with the weights of the single-process run over all 5 episodes."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if ws > 1:
        td.init_process_group("gloo", rank=rank, world_size=ws)
    from dhts_b200 import dist
    from dhts_b200.control import Trainer
    from itscp_env_cases import c4_env, c4_fixture
    env = c4_env(c4_fixture(), "cpu")
    lo, _ = dist.shard_range(5, rank, ws)

    def fake_rollout(action, differentiable, **kw):
        env.flags = type("F", (), {"check": staticmethod(lambda **k: None)})()
        R = action.shape[0]
        tgt = 0.3 + 0.05 * (lo + torch.arange(R, dtype=torch.float64)).unsqueeze(1)      # episode-dependent target
        return -((action.double() - tgt) ** 2).sum(1)
    env.rollout = fake_rollout
    torch.manual_seed(rank)                         # different initial weights per rank: the constructor must broadcast rank 0's
    if ws == 1:
        torch.manual_seed(0)
    tr = Trainer(env, lr=1e-2, tensorboard=False)
    losses = tr.train(5, 3, 1, 1, os.path.join(tmp, "ws%d" % ws))
    flat = torch.cat([p.detach().reshape(-1) for p in tr.controller.parameters()])
    q.put((rank, losses, flat[:64].tolist(), float(flat.double().sum())))
    if ws > 1:
        td.destroy_process_group()


def test_two_rank_trainer_matches_single_process(tmp_path):
    ctx = mp.get_context("spawn")
    out = {}
    for ws in (1, 2):
        port = 31500 + os.getpid() % 2000 + ws
        q = ctx.Queue()
        procs = [ctx.Process(target=_trainer_worker, args=(r, ws, port, q, str(tmp_path))) for r in range(ws)]
        for p in procs:
            p.start()
        out[ws] = sorted(q.get(timeout=240) for _ in range(ws))
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
    single = out[1][0]
    for rank, losses, head, total in out[2]:
        assert max(abs(a - b) for a, b in zip(losses, single[1])) < 1e-5 * max(1.0, abs(single[1][0]))      # fp32 controller, different summation order
        assert max(abs(a - b) for a, b in zip(head, single[2])) < 1e-5
        assert abs(total - single[3]) < 1e-4 * max(1.0, abs(single[3]))
    assert os.path.exists(os.path.join(str(tmp_path), "ws2", "eval.txt"))           # rank 0 alone writes the logs


def test_bench_inputs_do_not_depend_on_the_number_of_ranks():
    """bench.py's strong split: a lane's synthetic inputs are a function of (batch seed, global lane index) only, so N ranks
    working on contiguous lane blocks of ONE batch see exactly the rows of the unsharded batch (what
    `shard_equals_unshard_bitwise` then checks on the device results)."""
    import sys
    import types
    sys.path.insert(0, ROOT)
    import bench
    from dhts_b200 import dist
    a = types.SimpleNamespace(lanes=37, cells=16, micro_lanes=29, lane_vehicles=5, sim_steps=10)
    full = bench.make_arz_inputs(a, 0, torch)
    fidm = bench.make_idm_inputs(a, 0, torch)
    for ws in (2, 3, 8):
        rows = {k: [] for k in ("r0", "u0", "gr", "gu", "tr", "tu")}
        veh = {k: [] for k in ("p0", "v0", "tp", "tv")}
        par = []
        for r in range(ws):
            lo, hi = dist.shard_range(a.lanes, r, ws)
            part = bench.make_arz_inputs(a, 0, torch, lo, hi)
            for k in rows:
                rows[k].append(part[k])
            lo, hi = dist.shard_range(a.micro_lanes, r, ws)
            pi = bench.make_idm_inputs(a, 0, torch, lo, hi)
            assert pi["off"].tolist() == [i * a.lane_vehicles for i in range(hi - lo + 1)]
            for k in veh:
                veh[k].append(pi[k])
            par.append(pi["params"])
        for k in rows:
            assert torch.equal(torch.cat(rows[k]), full[k]), (ws, k)
        for k in veh:
            assert torch.equal(torch.cat(veh[k]), fidm[k]), (ws, k)
        assert torch.equal(torch.cat(par, dim=1), fidm["params"])
    # a different batch seed (the weak split: one batch per rank) gives different lanes
    assert not torch.equal(bench.make_arz_inputs(a, 1, torch)["r0"], full["r0"])
