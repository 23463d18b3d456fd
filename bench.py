#!/usr/bin/env python
"""Headline benchmark: fwd+bwd cell-updates/s (ARZ) and vehicle-updates/s (IDM) on the synthetic batch of
BASELINE.json configs[4]: per GPU 65536 independent ARZ lanes x 1024 cells + 4,194,304 IDM vehicles
(65536 lanes x 64), 1000 simulation steps forward + adjoint.

One bench "step" = one full pass of the hot path over that batch: ARZ rollout fwd + loss + adjoint, IDM
rollout fwd + loss + adjoint (one update = one cell / vehicle advanced one simulation step forward AND its
adjoint propagated one step back).  Lanes are independent, so ranks hold independent shards
(weak scaling) and the only collective is one all-reduce of the scalar losses per pass.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # our CUDA path
    python bench.py --impl reference ...                            # CPU arm (oracle port, all host threads)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is derived.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per update, single-pass streaming model, in scalars of the state dtype
ARZ_SCALARS_FWD, ARZ_SCALARS_BWD = 4, 6          # fwd: read (r,y) write (r,y); bwd: read (r,y), read adj, write adj
IDM_SCALARS_FWD, IDM_SCALARS_BWD = 10, 12        # fwd: read (p,v)+6 params, write (p,v); bwd: read 8 + adj 2 + write 2
SEED = 20221008


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--lanes", type=int, default=65536, help="ARZ lanes per GPU")
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--micro-lanes", type=int, default=65536, help="IDM lanes per GPU")
    ap.add_argument("--lane-vehicles", type=int, default=64)
    ap.add_argument("--sim-steps", type=int, default=1000)
    ap.add_argument("--ckpt-every", type=int, default=0,
                    help="ARZ checkpoint interval; 0 = auto: store every state (no recompute in the adjoint) and walk "
                         "the lanes in chunks that fit the free HBM, else 32 with segment recompute")
    ap.add_argument("--idm-ckpt-every", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-net", action="store_true", help="skip the secondary connected-network (ITSCP) measurement")
    ap.add_argument("--net-replicas", type=int, default=2048)
    return ap.parse_args()


def workload_name(a):
    return ("synthetic batch per GPU: %d ARZ lanes x %d cells + %d IDM vehicles (%d lanes x %d), %d steps fwd+bwd"
            % (a.lanes, a.cells, a.micro_lanes * a.lane_vehicles, a.micro_lanes, a.lane_vehicles, a.sim_steps))


# ----------------------------------------------------------------------------------------- synthetic inputs

def make_arz_inputs(a, shard, torch):
    """SURVEY 8(d) 'Synthetic ARZ input' (C1's physics, example/inverse/macro.py:48-49,88-89,246-252)."""
    g = torch.Generator().manual_seed(SEED + shard)
    B, N, umax = a.lanes, a.cells, 30.0
    r0 = torch.rand((B, N), generator=g, dtype=torch.float32)
    u0 = torch.rand((B, N), generator=g, dtype=torch.float32) * umax
    gr = torch.rand((B, 2), generator=g, dtype=torch.float32)
    gu = torch.rand((B, 2), generator=g, dtype=torch.float32) * umax
    tr = torch.rand((B, N), generator=g, dtype=torch.float32)
    tu = torch.rand((B, N), generator=g, dtype=torch.float32) * umax
    return dict(r0=r0, u0=u0, gr=gr, gu=gu, tr=tr, tu=tu, dx=5.0, umax=umax, dt=0.01)


def make_idm_inputs(a, shard, torch):
    """SURVEY 8(d) 'Synthetic IDM input' (example/inverse/micro.py:77-81; road/vehicle/micro_vehicle.py:88-109)."""
    g = torch.Generator().manual_seed(SEED + 7919 + shard)
    L, n, umax = a.micro_lanes, a.lane_vehicles, 30.0
    V = L * n
    rnd = lambda *s: torch.rand(s, generator=g, dtype=torch.float32)
    p0 = (torch.arange(n, dtype=torch.float32)[None, :] * 20.0 + rnd(L, n) * 10.0).reshape(V)
    v0 = 9.0 + 12.0 * rnd(V)
    params = torch.stack([(1.5 + 0.5 * rnd(V)) * umax, (1.0 + 0.5 * rnd(V)) * umax, (0.8 + 0.4 * rnd(V)) * umax,
                          1.0 + rnd(V), 0.2 + 0.4 * rnd(V), torch.full((V,), 5.0)])
    tp = p0 + 0.01 * 1000 * 15.0 + rnd(V)
    tv = 9.0 + 12.0 * rnd(V)
    off = (torch.arange(L + 1, dtype=torch.int64) * n).to(torch.int32)
    head = torch.tensor([[1000.0, 0.0]], dtype=torch.float32).repeat(L, 1)
    return dict(p0=p0, v0=v0, params=params, tp=tp, tv=tv, off=off, head=head, dt=0.01, n=n)


# ----------------------------------------------------------------------------------------- clocks sampler

class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 <= ts <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU arms

def cpu_port_rates(a, seconds_budget=12.0):
    """Times the oracle port (oracle/dhts_oracle.c, OpenMP over lanes) on a bounded sample of the workload:
    lanes of the config's shape (N cells / n vehicles), fewer lanes and fewer steps.  Returns rates + sample text."""
    import numpy as np
    from oracle import oracle as O
    cores = O.set_threads(len(os.sched_getaffinity(0)))     # explicit: torchrun exports OMP_NUM_THREADS=1
    rng = np.random.default_rng(SEED)
    N, n, umax = a.cells, a.lane_vehicles, 30.0

    def arz(B, T):
        r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 1, (B, N)) * umax
        gh = np.stack([rng.uniform(0, 1, (B, 2)), rng.uniform(0, 1, (B, 2)) * umax], -1)
        w = rng.normal(size=(B, N))
        t = time.perf_counter()
        O.arz_rollout(r0, u0, gh, 5.0, umax, 0.01, T, g_rT=w, g_uT=w / umax)
        return B * N * T / (time.perf_counter() - t)

    def idm(L, T):
        V = L * n
        p0 = (np.arange(n)[None] * 20.0 + rng.uniform(0, 10, (L, n))).ravel(); v0 = rng.uniform(9, 21, V)
        par = np.stack([rng.uniform(1.5, 2, V) * umax, rng.uniform(1, 1.5, V) * umax, rng.uniform(0.8, 1.2, V) * umax,
                        rng.uniform(1, 2, V), rng.uniform(0.2, 0.6, V), np.full(V, 5.0)])
        off = np.arange(L + 1) * n; head = np.tile([[1000.0, 0.0]], (L, 1)); w = rng.normal(size=V)
        t = time.perf_counter()
        O.idm_rollout(p0, v0, par, off, head, 0.01, T, g_pT=w, g_vT=w)
        return V * T / (time.perf_counter() - t)

    Ta, Ti = min(20, a.sim_steps), min(100, a.sim_steps)
    probe_a = arz(cores, Ta)                      # calibrate, then size the timed sample to the budget
    Ba = int(max(cores, min(a.lanes, probe_a * seconds_budget * 0.5 / (N * Ta))))
    rate_a = arz(Ba, Ta)
    probe_i = idm(cores * 4, Ti)
    Li = int(max(cores, min(a.micro_lanes, probe_i * seconds_budget * 0.5 / (n * Ti))))
    rate_i = idm(Li, Ti)
    sample = ("%d ARZ lanes x %d cells x %d steps fwd+bwd; %d IDM lanes x %d vehicles x %d steps fwd+bwd; "
              "OpenMP over lanes" % (Ba, N, Ta, Li, n, Ti))
    return rate_a, rate_i, cores, sample


def run_reference(a):
    """CPU arm.  The reference is pure Python (no C sources to compile into oracle/_ref) and cannot travel to the
    GPU box, so this times the oracle port with every host thread, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    budget = max(4.0, min(20.0, 150.0 / max(1, a.steps + a.warmup)))
    for _ in range(a.warmup):
        cpu_port_rates(a, budget)
    ra, ri, ms = [], [], []
    for _ in range(max(1, a.steps)):
        t = time.perf_counter()
        x, y, cores, sample = cpu_port_rates(a, budget)
        ms.append((time.perf_counter() - t) * 1e3); ra.append(x); ri.append(y)
    va, vi = sum(ra) / len(ra), sum(ri) / len(ri)
    line = {"impl": "reference", "metric": "fwd+bwd cell-updates/s", "value": va, "unit": "cell-updates/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sum(ms) / len(ms),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample_per_step": sample},
            "idm": {"value": vi, "unit": "vehicle-updates/s"},
            "cpu_baseline": {"value": va, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample,
                             "idm_value": vi, "idm_unit": "vehicle-updates/s",
                             "note": "C port of the reference step (oracle/); the reference's own Python path measured "
                                     "4.3e3 cell-updates/s and 9.3e3 vehicle-updates/s per core (BASELINE.md sec. 2)"},
            "e2e": {"value": va, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- connected network (secondary)

def network_bench(a, dev, dt_t, torch):
    """Secondary measurement: BASELINE.json configs[3]'s lane graph in macro mode -- 3 x 3 intersections,
    1 lane per road, 5 m access lanes (144 lanes, 288 cells), u_max 60, 30 Hz, 600 frames, 4 s signals -- rolled out
    for `net_replicas` candidate signal plans at once, forward + adjoint (gradient of the fused queue reward wrt
    every action), through dhts_net_rollout_{fwd,bwd}."""
    import numpy as np
    from dhts_b200.itscp import ItscpBatch, ItscpGrid
    grid = ItscpGrid(3, 1, 5.0, 5.0)
    env = ItscpBatch(grid, dev, speed_limit=60.0, simulation_frequency=30, signal_length=4.0, dtype=dt_t)
    R, T, L, NC = a.net_replicas, 600, grid.L, env.topo.NC
    rng = np.random.default_rng(SEED)
    nxt, route = env.topo.next, -np.ones((T, 2, L), dtype=np.int32)
    for t in range(T):          # RoadNetwork.create_random_macro_route (road_network.py:389-423)
        for l in rng.permutation(L):
            for n in (rng.permutation(nxt[l]) if nxt[l] else []):
                if route[t, 0, n] < 0:
                    route[t, 1, l] = n; route[t, 0, n] = l
                    break
    g = torch.Generator().manual_seed(SEED + 31)
    n_act = (T // env.frames_per_signal) * 9
    action = (0.1 + 0.8 * torch.rand((R, n_act), generator=g, dtype=torch.float32)).to(dt_t).to(dev).requires_grad_()
    sess = torch.rand((R, 5, L), generator=g, dtype=torch.float32).to(dt_t)          # itscp_random_schedule: 5 sessions
    incoming = sess.repeat_interleave(T // 5, dim=1).contiguous().to(dev)
    route_t = torch.tensor(route, device=dev)
    qk = torch.full((T,), 16.0 / 30.0, dtype=dt_t, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    fwd_ms, bwd_ms = [], []
    for it in range(3):
        action.grad = None
        e = [ev() for _ in range(4)]
        sig = grid.signals(action, T, env.frames_per_signal, soft=True)
        e[0].record()
        reward, states = env.rollout(action, incoming, route_t, T, differentiable=True, exact_constants=False, qk=qk)
        e[1].record()
        loss = reward.sum()
        e[2].record()
        loss.backward()
        e[3].record()
        torch.cuda.synchronize()
        if it:
            fwd_ms.append(e[0].elapsed_time(e[1])); bwd_ms.append(e[2].elapsed_time(e[3]))
        del states, reward, loss, sig
    f, b = sum(fwd_ms) / len(fwd_ms), sum(bwd_ms) / len(bwd_ms)
    upd = R * NC * T
    esz = 8 if dt_t == torch.float64 else 4
    return {"workload": "ITSCP 3x3 grid, macro mode: %d lanes / %d cells per replica, %d replicas, %d frames fwd+bwd "
                        "(signals from %d actions per replica, fused queue reward)" % (L, NC, R, T, n_act),
            "value": upd / ((f + b) / 1e3), "unit": "cell-updates/s", "fwd_ms": f, "bwd_ms": b,
            "replica_rollouts_per_s": R / ((f + b) / 1e3),
            "stored_state_bytes": (T + 1) * R * 3 * NC * esz, "gpu_launches": 2,
            "mean_reward": float(env.rollout(action.detach(), incoming, route_t, T, exact_constants=False, qk=qk)[0].mean()),
            "grad_abs_mean": float(action.grad.abs().mean()),
            "reference_note": "the reference steps this network lane by lane in Python: ~2.5-4.3e3 cell-updates/s per "
                              "core (BASELINE.md sec. 2), i.e. ~40-70 s per 600-frame episode fwd+bwd"}


def config4_bench(a, dev, dt_t, torch):
    """Secondary measurement: BASELINE.json configs[3] itself -- one training episode of run_itscp_hybrid.sh (hybrid mode,
    3 x 3 intersections, centre intersection micro, 600 frames, 45 actions, problem_1 inflow) through the headless env:
    action -> signals -> fused hybrid rollout -> queue reward with running-mean constants -> gradient wrt the actions.
    R = 1 is what the reference's trainer does per epoch (trainer.py:140-162 with 1 episode); R = 256 batches episodes."""
    import numpy as np
    from dhts_b200.itscp_env import ItscpEnv, problem_1
    np.random.seed(SEED)
    env = ItscpEnv(device=dev, dtype=dt_t)
    env.schedule_callback = problem_1
    env.config.update(num_intersection=3, lane_length=5.0, num_lane=1, policy_length=20, signal_length=4, mode="hybrid",
                      speed_limit=60.0)
    env.reset()
    g = torch.Generator().manual_seed(SEED + 41)
    out = {"workload": "run_itscp_hybrid.sh episode: 144 lanes (128 macro / 16 micro), 600 frames fwd+bwd, exact running-mean "
                       "queue reward, gradient wrt 45 actions"}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for R in (1, 256):
        action = (0.3 + 0.4 * torch.rand((R, env.action_size()), generator=g, dtype=torch.float32)).to(dt_t).to(dev).requires_grad_()
        env.resample_spawn_routes(R, g)
        ms = []
        for it in range(3):
            action.grad = None
            e0, e1 = ev(), ev()
            e0.record()
            reward = env.rollout(action, True)
            reward.sum().backward()
            e1.record()
            torch.cuda.synchronize()
            if it:
                ms.append(e0.elapsed_time(e1))
        bits, ncol = env.flags.read()
        out["R%d" % R] = {"ms_per_batch": sum(ms) / len(ms), "episodes_per_s": R / (sum(ms) / len(ms) / 1e3),
                          "mean_reward": float(reward.mean()), "grad_abs_mean": float(action.grad.abs().mean()),
                          "flag_bits": bits, "collision_steps": ncol}
    out["reference_note"] = ("the live reference needs ~200 s for the same episode forward + backward on one host core "
                             "(oracle/gen_golden_c4.py in the build container; it is single-threaded Python)")
    return out


# ----------------------------------------------------------------------------------------- our arm

def run_ours(a):
    import torch
    import torch.distributed as dist
    import dhts_b200
    from dhts_b200 import dist as shards
    from dhts_b200 import functional as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dhts_b200._lib.load()
    dt_t = torch.float64 if a.dtype == "f64" else torch.float32
    esz = 8 if a.dtype == "f64" else 4

    A = make_arz_inputs(a, rank, torch)
    M = make_idm_inputs(a, rank, torch)
    pin = lambda t: t.to(dt_t if t.is_floating_point() else t.dtype).pin_memory()
    hostA = {k: pin(v) for k, v in A.items() if hasattr(v, "shape")}
    hostM = {k: pin(v) for k, v in M.items() if hasattr(v, "shape")}
    devA = {k: v.to(dev) for k, v in hostA.items()}
    devM = {k: v.to(dev) for k, v in hostM.items()}
    flags = dhts_b200.Flags(dev)
    B, N, T = a.lanes, a.cells, a.sim_steps
    V = a.micro_lanes * a.lane_vehicles
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # ARZ plan: lanes per chunk and checkpoint interval (functional.arz_rollout_plan: every state stored when a
    # chunk of >= 296 lanes fits the free HBM, which removes the segment recompute from the adjoint)
    if a.ckpt_every > 0:
        arz_K = a.ckpt_every
        arz_chunk = F.arz_rollout_plan(B, N, T, dt_t, dev)[0] if arz_K == 1 else B
    else:
        arz_chunk, arz_K = F.arz_rollout_plan(B, N, T, dt_t, dev)
    arz_chunks = [(lo, min(B, lo + arz_chunk)) for lo in range(0, B, arz_chunk)]
    # one checkpoint arena for all chunks and passes (a chunk's checkpoints are dead once its adjoint has run)
    arz_arena = torch.empty(((T + arz_K - 1) // arz_K) * 2 * arz_chunk * N, dtype=dt_t, device=dev)

    def arz_pass(d, timers=None, before_chunk=None, after_chunk=None):
        """fwd + loss + adjoint of all B lanes, chunk of lanes by chunk of lanes (lanes are independent).
        before_chunk(i) / after_chunk(i, g_r0, g_u0) let the end-to-end pass pipeline its copies with the compute."""
        g_r0 = torch.empty_like(d["r0"]); g_u0 = torch.empty_like(d["u0"])
        total = torch.zeros((), dtype=dt_t, device=dev)
        for ci, (lo, hi) in enumerate(arz_chunks):
            if before_chunk: before_chunk(ci)
            r0 = d["r0"][lo:hi].detach().requires_grad_(); u0 = d["u0"][lo:hi].detach().requires_grad_()
            ev4 = [ev() for _ in range(4)] if timers is not None else None
            if ev4: ev4[0].record()
            rT, yT, uT = F.arz_rollout(r0, u0, d["gr"][lo:hi], d["gu"][lo:hi], A["dx"], A["umax"], A["dt"], T,
                                       ckpt_every=arz_K, flags=flags, ckpt_buffer=arz_arena)
            if ev4: ev4[1].record()
            loss = ((rT - d["tr"][lo:hi]) ** 2).sum() + ((uT - d["tu"][lo:hi]) ** 2).sum()   # example/inverse/macro.py:226-241
            if ev4: ev4[2].record()
            loss.backward()
            if ev4: ev4[3].record(); timers.append(ev4)
            g_r0[lo:hi] = r0.grad; g_u0[lo:hi] = u0.grad
            total += loss.detach()
            if after_chunk: after_chunk(ci, g_r0, g_u0)
            del rT, yT, uT, loss, r0, u0
        return total, g_r0, g_u0

    def idm_pass(d, timers=None):
        p0 = d["p0"].detach().requires_grad_(); v0 = d["v0"].detach().requires_grad_()
        ev4 = [ev() for _ in range(4)] if timers is not None else None
        if ev4: ev4[0].record()
        pT, vT = F.idm_rollout(p0, v0, d["params"], d["off"], d["head"], M["dt"], T, ckpt_every=a.idm_ckpt_every,
                               flags=flags, max_lane=M["n"])
        if ev4: ev4[1].record()
        loss = ((pT - d["tp"]) ** 2).sum() + ((vT - d["tv"]) ** 2).sum()     # example/inverse/micro.py:221-236
        if ev4: ev4[2].record()
        loss.backward()
        if ev4: ev4[3].record(); timers.append(ev4)
        return loss.detach(), p0.grad, v0.grad

    def reduce_loss(la, li):
        return shards.reduce_losses(la, li)     # the only collective: scalar losses over the lane shards

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident pass (value): inputs already in HBM
    for _ in range(a.warmup):
        la, _, _ = arz_pass(devA); li, _, _ = idm_pass(devM); reduce_loss(la, li)
    barrier()
    clocks = Clocks(local) if rank == 0 else None
    t_wall0 = time.time()
    tA, tM = [], []              # CUDA events (start, fwd done, loss done, bwd done) per launch pair
    e0, e1 = ev(), ev()
    e0.record()
    for s in range(a.steps):
        la, _, _ = arz_pass(devA, tA); li, _, _ = idm_pass(devM, tM); losses = reduce_loss(la, li)
    e1.record()
    barrier()
    t_wall1 = time.time()
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    total_ms = e0.elapsed_time(e1)
    arz_fwd = sum(t[0].elapsed_time(t[1]) for t in tA); arz_bwd = sum(t[2].elapsed_time(t[3]) for t in tA)
    arz_all = sum(t[0].elapsed_time(t[3]) for t in tA)
    idm_fwd = sum(t[0].elapsed_time(t[1]) for t in tM); idm_bwd = sum(t[2].elapsed_time(t[3]) for t in tM)
    idm_all = sum(t[0].elapsed_time(t[3]) for t in tM)
    bits, ncol = flags.read()

    # ---- end-to-end pass: host (pinned) inputs -> H2D -> rollouts -> D2H of losses and gradients, every step
    e2e_ms = None
    h2d = d2h = 0
    if not a.no_e2e:
        outA = [torch.empty((B, N), dtype=dt_t).pin_memory() for _ in range(2)]
        outM = [torch.empty((V,), dtype=dt_t).pin_memory() for _ in range(2)]
        h2d = sum(v.numel() * v.element_size() for v in hostA.values()) + sum(v.numel() * v.element_size() for v in hostM.values())
        d2h = sum(t.numel() * t.element_size() for t in outA + outM) + 2 * esz

        # Copies are pipelined with the compute, lane chunk by lane chunk: inputs of chunk i+1.. travel on a copy-in
        # stream while chunk i computes; its gradients leave on a copy-out stream.  Everything is inside the timed
        # region; the pass ends with the host read of the reduced losses after both streams have drained.
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        dA = {k: torch.empty_like(v, device=dev) for k, v in hostA.items()}
        dM = {k: torch.empty_like(v, device=dev) for k, v in hostM.items()}
        pass_done = ev()
        pass_done.record()

        def e2e_pass():
            cur = torch.cuda.current_stream(dev)
            in_ready = []
            s_in.wait_event(pass_done)                      # the previous pass no longer reads the staging buffers
            with torch.cuda.stream(s_in):
                for lo, hi in arz_chunks:
                    for k, v in hostA.items():
                        dA[k][lo:hi].copy_(v[lo:hi], non_blocking=True)
                    e = ev(); e.record(s_in); in_ready.append(e)
                for k, v in hostM.items():
                    dM[k].copy_(v, non_blocking=True)
                idm_ready = ev(); idm_ready.record(s_in)

            def before(ci):
                cur.wait_event(in_ready[ci])

            def after(ci, g_r0, g_u0):
                lo, hi = arz_chunks[ci]
                done = ev(); done.record(cur)
                s_out.wait_event(done)
                with torch.cuda.stream(s_out):
                    outA[0][lo:hi].copy_(g_r0[lo:hi], non_blocking=True); outA[1][lo:hi].copy_(g_u0[lo:hi], non_blocking=True)
                g_r0.record_stream(s_out); g_u0.record_stream(s_out)

            la, _, _ = arz_pass(dA, before_chunk=before, after_chunk=after)
            cur.wait_event(idm_ready)
            li, gp0, gv0 = idm_pass(dM)
            outM[0].copy_(gp0, non_blocking=True); outM[1].copy_(gv0, non_blocking=True)
            cur.wait_stream(s_out)
            pass_done.record(cur)
            return reduce_loss(la, li).cpu()

        e2e_pass()
        barrier()
        f0, f1 = ev(), ev()
        f0.record()
        for _ in range(a.steps):
            e2e_pass()
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1)

    # ---- max over ranks
    times = torch.tensor([total_ms, arz_all, idm_all, arz_fwd, arz_bwd, idm_fwd, idm_bwd, e2e_ms or 0.0],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, arz_all, idm_all, arz_fwd, arz_bwd, idm_fwd, idm_bwd, e2e_ms_max = times.tolist()

    if rank == 0:
        K = a.steps
        cell_updates = world * B * N * T * K
        veh_updates = world * V * T * K
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"
        # dominant kernel: arz_rollout_bwd (one launch per pass per GPU)
        nl = len(arz_chunks)                                   # adjoint launches per pass (one per lane chunk)
        bwd_bytes = B * N * T * ARZ_SCALARS_BWD * esz / nl      # algorithmic bytes of ONE launch (average chunk)
        bwd_s = arz_bwd / K / nl / 1e3                         # average duration of one launch
        # DRAM bytes of the same kernel from the committed `ncu --set full` capture (profiles/traffic.json, written
        # by scripts/ncu_summary.py): captured on a smaller batch with the same checkpoint interval, so it is
        # carried per cell-step and scaled to this launch
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            tj = tj.get("arz_rollout_bwd_%s_k%d" % (a.dtype, arz_K)) or tj["arz_rollout_bwd_" + a.dtype]
            traffic = tj["dram_bytes_per_cell_step"] * B * N * T / len(arz_chunks)
        except Exception:
            pass
        both_bytes = B * N * T * (ARZ_SCALARS_FWD + ARZ_SCALARS_BWD) * esz
        both_s = (arz_fwd + arz_bwd) / K / 1e3
        idm_bytes = V * T * (IDM_SCALARS_FWD + IDM_SCALARS_BWD) * esz
        idm_s = (idm_fwd + idm_bwd) / K / 1e3
        line = {
            "metric": "fwd+bwd cell-updates/s", "value": cell_updates / (arz_all / 1e3), "unit": "cell-updates/s",
            "n_gpus": world, "steps": K, "warmup": a.warmup, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "arz": {"lanes_per_gpu": B, "cells": N, "dx": 5.0, "dt": 0.01,
                                                             "u_max": 30.0},
                       "idm": {"vehicles_per_gpu": V, "lane_vehicles": a.lane_vehicles, "params": "per-vehicle"},
                       "sim_steps": T, "ckpt_every": arz_K, "arz_lanes_per_chunk": arz_chunk,
                       "arz_chunks_per_pass": len(arz_chunks), "idm_ckpt_every": a.idm_ckpt_every, "parallelism": "lane shards x%d, all-reduce(loss)" % world,
                       "l2": "inputs (%.1f GB per pass) are larger than the 126 MB L2" % ((2 * B * N + 8 * V) * esz / 1e9)},
            "idm": {"metric": "fwd+bwd vehicle-updates/s", "value": veh_updates / (idm_all / 1e3),
                    "unit": "vehicle-updates/s", "ms_per_step": idm_all / K},
            "phase_ms_per_step": {"arz_fwd": arz_fwd / K, "arz_bwd": arz_bwd / K, "arz_total": arz_all / K,
                                  "idm_fwd": idm_fwd / K, "idm_bwd": idm_bwd / K, "idm_total": idm_all / K},
            "gpu_launches": (2 * len(arz_chunks) + 2) * K,
            "roofline": {"kernel": "arz_rollout_bwd_reg_kernel<%s, 4, 2, %d>" % ("double" if a.dtype == "f64" else "float",
                                                                               0 if arz_K == 1 else 2),
                         "bound": "hbm", "achieved": bwd_bytes / bwd_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bwd_bytes / bwd_s / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bwd_bytes,
                         "launches_per_pass": len(arz_chunks),
                         "note": "achieved = 6 scalars x cell-steps of one adjoint launch / its mean CUDA-event time; "
                                 "traffic = measured DRAM bytes of that launch (ncu, scaled per cell-step)"},
            "roofline_fwd_bwd": {"arz": {"achieved": both_bytes / both_s / 1e9, "frac": both_bytes / both_s / 1e9 / peak,
                                         "bytes_per_update": (ARZ_SCALARS_FWD + ARZ_SCALARS_BWD) * esz},
                                 "idm": {"achieved": idm_bytes / idm_s / 1e9, "frac": idm_bytes / idm_s / 1e9 / peak,
                                         "bytes_per_update": (IDM_SCALARS_FWD + IDM_SCALARS_BWD) * esz},
                                 "unit": "GB/s", "peak": peak},
            "clocks": clk,
            "flags": {"bits": bits, "collisions": ncol},
            "losses": [float(x) for x in losses.tolist()],
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": cell_updates / (e2e_ms_max / 1e3), "unit": "cell-updates/s",
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max / K,
                           "note": "whole pass (ARZ + IDM) from pinned host buffers, gradients and losses read back, copies "
                                   "pipelined with the compute per lane chunk; value counts ARZ cell-updates over the "
                                   "WHOLE pass time (IDM and copies included)"}
        if not a.no_net:
            del arz_arena, devA, devM
            torch.cuda.empty_cache()
            line["itscp_net"] = network_bench(a, dev, dt_t, torch)
            torch.cuda.empty_cache()
            line["itscp_c4"] = config4_bench(a, dev, dt_t, torch)
        if not a.no_cpu_baseline and world == 1:      # the CPU leg is reported at N = 1 only
            ra, ri, cores, sample = cpu_port_rates(a)
            line["cpu_baseline"] = {"value": ra, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                    "sample": sample, "idm_value": ri, "idm_unit": "vehicle-updates/s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
