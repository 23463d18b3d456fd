"""Headless ITSCP environment: ``example/control/itscp/_env.py`` without highway-env / gym /
pygame, stepping the fused network rollouts instead of lane objects.

Same surface as the reference's ``ItscpEnv`` for everything the trainer and ``run.py`` touch
(example/control/itscp/run.py:48-59, example/control/trainer.py:22-196):

    env = ItscpEnv(); env.schedule_callback = problem_1; env.config['mode'] = 'hybrid'; ...; env.reset()
    obs = env.observe(); obs, reward, terminal, info = env.step(action, differentiable)

``step`` runs one whole policy (``policy_length`` seconds = all frames, _env.py:744-755) in ONE kernel launch and
returns the queue-length reward (_env.py:618-648,770-797) as a 0-dim tensor whose ``backward`` reaches ``action``
through the adjoint kernels.  ``rollout`` is the batched form: R candidate actions / inflow schedules / spawn-route
draws at once (episodes of an epoch, scenario replicas of a multi-GPU run).

Modes: ``macro`` (every lane a dMacroLane -> ``network.net_rollout``), ``hybrid`` (lanes of interior
intersections are dMicroLanes, _env.py:490-500 -> ``hybrid_network.hybrid_rollout``) and ``micro`` (every lane a
MicroLane, _env.py:484-488; ``run_itscp_micro.sh``): the same hybrid rollout over an all-micro network whose boundary
lanes are fed from the waiting lists ``_make_micro_route`` fills (_env.py:202-219) by the stochastic rule of
``ItscpRoadNetwork.setup_micro_boundary`` (_simulator.py:153-174).  The uniform draws of that rule are taken from
``np.random`` in the reference's order and count, so a seeded episode -- and the ``np.random`` state it leaves
behind -- match the reference's.  Rendering does not exist: ``info['img']`` is a list of ``None`` (what the
reference returns when ``render_eval`` is off, _env.py:728-742).
"""
from __future__ import annotations

import copy
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .hybrid_network import HybridNetTopology, hybrid_rollout
from .itscp import ItscpGrid, LaneInfo
from .network import MODE_ITSCP, net_rollout

LaneID = LaneInfo       # _env.py:23-60: the schedule callbacks read .row .col .loc .ploc .approaching .lane_id

default_config = {      # example/control/itscp/_env_config.py:1-84 (viewer keys dropped)
    "num_intersection": 1, "num_lane": 3, "lane_length": 20, "speed_limit": 60, "cell_length": 5, "vehicle_length": 5,
    "simulation_frequency": 30, "policy_length": 10, "signal_length": 2, "action_min": 0.1, "action_max": 0.9,
    "duration": 1, "static_speed": 0.2, "num_schedule_obs": 10, "max_num_micro_vehicle_per_lane": 10, "mode": "macro",
    "render": False, "random_seed": 0,
    # headless extras
    "veh_cap": 8,            # vehicle slots per micro lane in the fused hybrid kernel
    "max_spawn": 64,         # spawn-route draws per micro lane and episode
}


class Box:
    """The two attributes of ``gym.spaces.Box`` the trainer reads (trainer.py:27-33,179-184)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low, dtype=dtype); high = np.asarray(high, dtype=dtype)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=dtype); high = np.full(shape, high, dtype=dtype)
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


def itscp_random_schedule(lane_id: List[LaneID], num_timestep: int):
    """_env.py:62-92: five sessions of constant random inflow per lane."""
    num_session = 5
    per = num_timestep // num_session
    schedule: Dict[LaneID, List[float]] = {}
    for id in lane_id:
        cur: List[float] = []
        for _ in range(num_session):
            r = np.random.random((1)).item()
            cur.extend([r] * per)
            cur = cur[:num_timestep]
        schedule[id] = cur
    return schedule


def problem(lane_id: List[LaneID], num_timestep: int, num_session: int):
    """example/control/itscp/problem.py:5-68: per session one of N-S / W-E carries the heavy inflow."""
    per = num_timestep // num_session
    schedule: Dict[LaneID, List[float]] = {}
    session_dir: List[str] = []
    for i in range(num_session):
        if i == 0:
            session_dir.append("NS" if np.random.random((1)).item() > 0.5 else "WE")
        else:
            session_dir.append("WE" if session_dir[-1] == "NS" else "NS")
    for id in lane_id:
        cur: List[float] = []
        for s in range(num_session):
            r = np.random.random((1)).item()
            if session_dir[s] == "NS":
                hot = id.loc in ("north", "south")
            else:
                hot = id.loc in ("west", "east")
            r = 0.9 + r * 0.1 if hot else 0.0 + r * 0.01
            cur.extend([r] * per)
        schedule[id] = cur[:num_timestep]
    return schedule


def problem_1(lane_id, num_timestep):
    return problem(lane_id, num_timestep, 1)


def problem_2(lane_id, num_timestep):
    return problem(lane_id, num_timestep, 2)


def problem_3(lane_id, num_timestep):
    return problem(lane_id, num_timestep, 3)


def running_mean_constants(samples: torch.Tensor, valid: Optional[torch.Tensor], scale: float = 16.0,
                           window: int = 100_000) -> torch.Tensor:
    """``scale / |RunningMean.mean()|`` after appending each VALID sample in order (common/rms.py:3-22 as used by
    _env.py:557-575,697-702): the mean covers the last `window` valid samples up to and including the current one.
    samples / valid are [M] or [R, M] (R independent sequences, one RunningMean each); entries at invalid positions are
    meaningless.  Accumulates in float64 (the reference in float32)."""
    shape = samples.shape
    d = samples.detach().to(torch.float64).reshape(-1, shape[-1]) if samples.dim() > 1 else samples.detach().to(torch.float64).reshape(1, -1)
    m = torch.ones_like(d, dtype=torch.bool) if valid is None else valid.reshape(d.shape)
    d = torch.where(m, d, torch.zeros_like(d))
    cs = torch.cumsum(d, 1)
    n = torch.cumsum(m.to(torch.int64), 1)
    cnt = torch.clamp(n, max=window)
    over = n - window                                    # how many valid samples have left the window
    if int(n[:, -1].max()) > window:
        first = torch.searchsorted(n, torch.clamp(over, min=1))      # position of the `over`-th valid sample, row by row
        drop = torch.where(over > 0, torch.gather(cs, 1, torch.clamp(first, max=d.shape[1] - 1)), torch.zeros_like(cs))
    else:
        drop = torch.zeros_like(cs)
    mean = (cs - drop) / torch.clamp(cnt, min=1).to(torch.float64)
    return (scale / mean.abs()).reshape(shape)


class ItscpEnv:
    """Intersection signal control problem, headless (see module docstring).  Not a gym env: the reference subclasses
    highway-env's AbstractEnv only for its viewer and config plumbing."""

    def __init__(self, schedule_callback: Callable = itscp_random_schedule, device=None, dtype=torch.float64):
        self.schedule_callback = schedule_callback
        self.config = dict(default_config)
        self.device = torch.device(device) if device is not None else None
        self.dtype = dtype
        self.render_eval = False
        self.grid: Optional[ItscpGrid] = None
        self.lane: Dict[LaneID, int] = {}
        self.schedule: Dict[LaneID, List[float]] = {}
        self.steps = self.time = 0
        self.done = False
        self.flags: Optional[_lib.Flags] = None
        self.last: Dict[str, object] = {}

    @classmethod
    def default_config(cls) -> dict:
        return dict(default_config)

    def __deepcopy__(self, memo):
        """trainer.py:172 deep-copies the env once per episode to start from a clean simulator.  The simulator state here
        lives in the kernels' buffers of one ``step`` call, so a copy shares topology and schedule and resets the counters."""
        c = copy.copy(self)
        c.config = dict(self.config)
        c.steps = c.time = 0
        c.done = False
        c.last = {}
        return c

    # ------------------------------------------------------------------ construction (_env.py:143-223)
    def action_size(self) -> int:
        simulation_length = self.config["policy_length"] * self.config["duration"]
        return int(simulation_length / self.config["signal_length"]) * (self.num_intersection ** 2)

    def reset(self):
        return self._reset()

    def _reset(self):
        cfg = self.config
        if cfg["random_seed"] > 0:
            np.random.seed(cfg["random_seed"])
        if cfg["mode"] not in ("macro", "micro", "hybrid"):
            raise ValueError("mode must be 'macro', 'micro' or 'hybrid' (_env.py:480-500)")
        if self.device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("ItscpEnv steps CUDA kernels only (no CPU fallback)")
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_intersection = int(cfg["num_intersection"])
        self.num_lane = int(cfg["num_lane"])
        self.num_timestep = int(cfg["policy_length"] * cfg["duration"] * cfg["simulation_frequency"])
        self._make_road()
        ids = list(self.lane.keys())
        self.schedule = self.schedule_callback(ids, self.num_timestep)
        self.observation_space = Box(0, 1, shape=(cfg["num_schedule_obs"] * len(self.lane),), dtype=np.float32)
        self.action_space = Box([cfg["action_min"]] * self.action_size(), [cfg["action_max"]] * self.action_size())
        self.time = self.steps = 0
        self.done = False
        self.reward_queue_c = -1.0
        self._make_macro_route()
        self._make_micro_route()
        self._incoming = None
        return self.observe()

    def _make_road(self):
        cfg = self.config
        g = ItscpGrid(self.num_intersection, self.num_lane, float(cfg["lane_length"]), float(cfg["cell_length"]))
        self.grid = g
        self.lane = {info: i for i, info in enumerate(g.lanes)}
        n = self.num_intersection
        if cfg["mode"] == "hybrid":      # _env.py:490-500
            self.kind = [0 if (i.row == 0 or i.row == n - 1 or i.col == 0 or i.col == n - 1) else 1 for i in g.lanes]
        elif cfg["mode"] == "micro":     # _env.py:484-488
            self.kind = [1] * g.L
        else:
            self.kind = [0] * g.L
        self.hybrid = any(self.kind)
        self.micro_mode = cfg["mode"] == "micro"
        self.veh_cap = int(cfg["veh_cap"])
        if self.micro_mode:              # a lane of length len holds at most len / vehicle_length + 1 vehicles bumper to bumper
            self.veh_cap = max(self.veh_cap, int(max(l.length for l in g.lanes) / float(cfg["vehicle_length"])) + 3)
        if self.hybrid:
            self.topo = HybridNetTopology(self.kind, g.num_cell, g.dx, [l.length for l in g.lanes], g.links, self.device,
                                          MODE_ITSCP, veh_cap=self.veh_cap, veh_len=float(cfg["vehicle_length"]),
                                          sources=self.micro_mode, enumerate_routes=not self.micro_mode)
        else:
            self.topo = g.topology(self.device)
        self.boundary = g.boundary_lanes()
        # sample layout of the queue reward: lanes in id order, cells of a macro lane / vehicle slots of a micro lane
        # (tail first, _env.py:676-716)
        t = self.topo
        order, is_cell = [], []
        cap = self.veh_cap if self.hybrid else 0
        cell_off = t.cell_off
        mic = 0
        for l in range(g.L):
            if self.kind[l]:
                order.extend(t.NC + mic * cap + j for j in range(cap)); is_cell.extend([False] * cap); mic += 1
            else:
                order.extend(range(cell_off[l], cell_off[l + 1])); is_cell.extend([True] * (cell_off[l + 1] - cell_off[l]))
        self._order = torch.tensor(order, dtype=torch.long, device=self.device)
        self._inv_order = torch.empty_like(self._order)
        self._inv_order[self._order] = torch.arange(len(order), device=self.device)

    def _make_macro_route(self):
        """_env.py:194-200: one random MacroRoute per frame (road_network.py:389-423, same np.random draws)."""
        g, T = self.grid, self.num_timestep
        nxt = self.topo.next
        tab = -np.ones((T, 2, g.L), dtype=np.int32)
        ids = list(range(g.L))
        for t in range(T):
            for lane_id in np.random.permutation(ids):
                if self.kind[lane_id]:
                    continue
                for nid in np.random.permutation(nxt[lane_id]):
                    if tab[t, 0, nid] == -1:
                        tab[t, 1, lane_id] = nid
                        tab[t, 0, nid] = lane_id
                        break
        self.macro_route_schedule = torch.tensor(tab, dtype=torch.int32, device=self.device)

    def _make_micro_route(self):
        """_env.py:202-219: ``max_num_micro_vehicle_per_lane`` waiting vehicles with random routes for EVERY lane, in every
        mode (same ``np.random.randint`` draws as ``create_random_route``, road_network.py:604-646).  Only micro mode reads
        them (boundary lanes, _simulator.py:153-174: popped from the END of the list).  In hybrid mode vehicles enter micro
        lanes through macro->micro spawning instead (every micro lane has a predecessor); each spawn
        draws a random route (conversion.py:53-57): R sets of ``max_spawn`` routes per micro lane, re-drawn by
        ``resample_spawn_routes``."""
        self._spawn_routes = None
        K = int(self.config["max_num_micro_vehicle_per_lane"])
        nxt = self.topo.next
        self.waiting_route: List[List[List[int]]] = []
        for lane_id in range(self.grid.L):
            rows = []
            for _ in range(K):
                route, cur = [], lane_id
                for _ in range(32):                              # MAX_ROUTE_LENGTH, road_network.py:15
                    route.append(cur)
                    cand = nxt[cur]
                    if not cand:
                        break
                    i = i0 = int(np.random.randint(0, len(cand)))
                    while cand[i] in route:
                        i = (i + 1) % len(cand)
                        if i == i0:
                            break
                    cur = cand[i]
                rows.append(route)
            self.waiting_route.append(rows)
        if self.micro_mode:       # pop order: the k-th vehicle that enters lane l rides waiting_route[l][K - 1 - k]
            ids = [[self.topo.route_id(r) for r in reversed(rows)] for rows in self.waiting_route]
            self._wait_ids = torch.tensor(ids, dtype=torch.int32, device=self.device).reshape(self.grid.L, K)

    def resample_spawn_routes(self, R: int, generator: Optional[torch.Generator] = None):
        if self.hybrid:
            self._spawn_routes = self.topo.random_spawn_routes(R, int(self.config["max_spawn"]), generator)
        return self._spawn_routes

    # ------------------------------------------------------------------ observation (_env.py:517-535)
    def observe(self):
        obs = []
        k_obs = self.config["num_schedule_obs"]
        has_prev = [len(p) > 0 for p in self.topo.prev]
        for lane_id, idx in self.lane.items():
            sc = self.schedule[lane_id]
            t = len(sc) // k_obs
            for k in range(k_obs):
                if not has_prev[idx]:
                    t0, t1 = int(t * k), min(int(t * k + t), len(sc))
                    obs.append(sum(sc[t0:t1]) / (t1 - t0))
                else:
                    obs.append(0)
        return np.array(obs).astype(self.observation_space.dtype)

    def incoming(self) -> torch.Tensor:
        """[T, L] scheduled inflow density of the boundary lanes (_env.py:605-616)."""
        if self._incoming is None:
            inc = np.zeros((self.num_timestep, self.grid.L))      # the reference stores -1 for lanes it never reads
            for lane_id, idx in self.lane.items():
                if idx in self.boundary:
                    inc[:, idx] = np.asarray(self.schedule[lane_id][:self.num_timestep], dtype=np.float64)
            self._incoming = torch.tensor(inc, dtype=self.dtype, device=self.device)
        return self._incoming

    # ------------------------------------------------------------------ signals (_env.py:885-962)
    def lane_signal_info(self, lane_id: LaneID, action, curr_frame: int, differentiable: bool):
        cfg = self.config
        fps = cfg["simulation_frequency"] * cfg["signal_length"]
        n2 = self.num_intersection ** 2
        phase = min(curr_frame // fps, len(action) // n2 - 1)
        a = torch.as_tensor(action[int(phase) * n2 + lane_id.row * self.num_intersection + lane_id.col])
        progress = min((curr_frame % fps) / fps, 1.0)
        we = (lane_id.ploc if lane_id.loc == "mid" else lane_id.loc) in ("west", "east")
        d = (a - progress) if we else (progress - a)
        s = torch.sigmoid(torch.clamp(d * 32.0, -16.0, 16.0)) if differentiable else float(d > 0)
        if lane_id.loc == "mid":
            return s, 1.0
        if not lane_id.approaching:
            return 1.0, 1.0
        return 1.0, s

    # ------------------------------------------------------------------ simulation + reward
    def rollout(self, action: torch.Tensor, differentiable: bool, incoming: Optional[torch.Tensor] = None,
                spawn_routes: Optional[torch.Tensor] = None, keep_states: bool = False,
                src_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
        """R episodes at once.  action [R, A] (any device / float dtype; gradients flow back to it); incoming [R, T, L]
        (default: this env's schedule for every replica); spawn_routes [R, ML, KS] or [ML, KS]; src_rand [n] or [R, n]
        (micro mode): the uniform draws of the waiting-list sources -- default: taken from ``np.random`` exactly as the
        reference's episode would (replica after replica).  Returns reward [R]."""
        cfg = self.config
        dev, dtype, T = self.device, self.dtype, self.num_timestep
        act = action.to(device=dev, dtype=dtype)
        R = act.shape[0]
        umax, dt = float(cfg["speed_limit"]), 1.0 / cfg["simulation_frequency"]
        fps = int(cfg["simulation_frequency"] * cfg["signal_length"])
        sig = self.grid.signals(act, T, fps, soft=differentiable)
        inc = self.incoming().unsqueeze(0).expand(R, -1, -1) if incoming is None else incoming.to(device=dev, dtype=dtype)
        topo = self.topo
        r0 = torch.zeros((R, topo.NC), dtype=dtype, device=dev)          # lanes start empty: ARZ.FullQ(u_max), _arz.py:59-63
        u0 = torch.full((R, topo.NC), umax, dtype=dtype, device=dev)
        self.flags = _lib.Flags(dev)
        static = float(cfg["static_speed"])
        veh_len = float(cfg["vehicle_length"])
        w = (topo.real("dx", dtype) if self.hybrid else topo.dx(dtype))[topo.lane_of_cell()] / veh_len
        if self.hybrid:
            rng_state = None
            if self.micro_mode:
                spawn_routes = self._wait_ids if spawn_routes is None else spawn_routes
                if src_rand is None:      # one draw per boundary lane and frame at most; the unused tail is handed back below
                    n_max = T * sum(topo.src)
                    rng_state = np.random.get_state() if R == 1 else None
                    src_rand = torch.tensor(np.random.random((R, n_max)), dtype=dtype, device=dev)
            elif spawn_routes is None:
                if self._spawn_routes is None or self._spawn_routes.shape[0] != R:
                    self.resample_spawn_routes(R)
                spawn_routes = self._spawn_routes
            st = hybrid_rollout(topo, r0, u0, umax, dt, T, sig=sig, incoming=inc.contiguous(), route=self.macro_route_schedule,
                                spawn_route=spawn_routes, soft=differentiable, flags=self.flags,
                                src_rand=src_rand.to(device=dev, dtype=dtype) if self.micro_mode else None)
            if rng_state is not None:     # leave np.random where the reference's episode would have left it
                used = int(st.aux[-1, 0, topo.A_DRAW].detach().round().item())
                np.random.set_state(rng_state)
                if used:
                    np.random.random((used,))
                self.last_draws = used
            cells = st.cells
            r = cells[1:, :, 0].transpose(0, 1); u = cells[1:, :, 2].transpose(0, 1)          # [R, T, NC]
            _, v, _, valid = st.by_rank()
            cap = topo.veh_cap
            assert cap == self.veh_cap
            v = v[1:].transpose(0, 1); valid = valid[1:].transpose(0, 1)                      # [R, T, ML, cap] head first
            cnt = st.count[1:].transpose(0, 1).unsqueeze(-1)                                  # [R, T, ML, 1]
            slot = torch.arange(cap, device=dev)
            tail_idx = torch.clamp(cnt - 1 - slot, min=0)                                     # tail first
            v_tail = torch.gather(v, 3, tail_idx); ok_tail = slot < cnt
            speed = torch.cat([u, v_tail.reshape(R, T, -1)], dim=2)
            mask = torch.cat([torch.ones_like(u, dtype=torch.bool), ok_tail.reshape(R, T, -1)], dim=2)
            weight = torch.cat([r * w, ok_tail.reshape(R, T, -1).to(dtype)], dim=2)
            lane_of = torch.cat([topo.lane_of_cell(), torch.tensor(topo.micro, device=dev).repeat_interleave(cap)])
        else:
            states, _ = net_rollout(topo, r0, u0, umax, dt, T, sig=sig, incoming=inc.contiguous(),
                                    route=self.macro_route_schedule, soft=differentiable, veh_len=veh_len,
                                    static_speed=static, flags=self.flags)
            r = states[1:, :, 0].transpose(0, 1); u = states[1:, :, 2].transpose(0, 1)
            speed, weight = u, r * w
            mask = torch.ones_like(u, dtype=torch.bool)
            lane_of = topo.lane_of_cell()
            st = states
        d = static - speed                                                                    # [R, T, S]
        if differentiable:       # _env.py:557-575: sigmoid with the running-mean constant, samples in lane-id order
            do, mo = d[..., self._order], mask[..., self._order]
            k = running_mean_constants(do.reshape(R, -1), mo.reshape(R, -1)).reshape(R, T, -1)      # one RunningMean per episode
            k = k[..., self._inv_order].to(dtype)
            is_static = torch.sigmoid(torch.clamp(d * k, -16.0, 16.0))
        else:                    # _env.py:576-586
            is_static = (d > 0).to(dtype)
        per = torch.where(mask, is_static * weight, torch.zeros_like(weight))
        q = torch.zeros((R, T, self.grid.L), dtype=dtype, device=dev).index_add(2, lane_of, per)
        reward = self.reward_queue_c * (q ** 2.0).sum(dim=(1, 2)) * dt
        self.last = {"states": st} if keep_states else {}
        self.time = T
        return reward

    def step(self, action, differentiable: bool):
        """_env.py:537-555: (obs, reward, terminal, info); reward is 0-dim and differentiable wrt `action`."""
        self.steps += 1
        a = torch.as_tensor(action)
        reward = self.rollout(a.reshape(1, -1), differentiable)[0]
        self.flags.check(quiet_collisions=True)
        obs = self.observe()
        terminal = self.steps >= self.config["duration"] or self.time >= self.num_timestep
        info = {"img": [None] * self.num_timestep}
        return obs, reward, terminal, info
