"""Headless intersection-signal-control (ITSCP) scenario over the fused network rollout.

The reference's ITSCP environment (example/control/itscp/_env.py) needs highway-env, gym and pygame only
for rendering and bookkeeping; what the signal optimiser differentiates through is
  * the lane graph of the n x n grid of 4-way intersections            _env.py:233-439
  * lane_signal_info: signal of every lane at every frame from the action vector   _env.py:885-962
  * ItscpRoadNetwork's signal-blended ghost cells + the macro lane step  _simulator.py:56-142
  * the queue-length reward with its running-mean sigmoid constant     _env.py:557-742,770-797; common/rms.py
This module restates those pieces for MACRO mode on top of ``dhts_b200.network.net_rollout``: many candidate
signal plans / inflow schedules (replicas) are rolled out and differentiated in one launch.

Lane numbering, link order and random-route draws follow the reference's creation order so that seeded
runs agree lane by lane with an ``ItscpRoadNetwork`` built by ``_make_road``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .network import MODE_ITSCP, MacroNetTopology, net_rollout

LANE_WIDTH = 4.0          # highway_env AbstractLane.DEFAULT_WIDTH, _env.py:233


@dataclass(eq=False)     # identity hash: schedule callbacks key dicts by lane id objects (_env.py:23-60)
class LaneInfo:
    row: int
    col: int
    loc: str                 # south / west / north / east / mid
    ploc: Optional[str]      # approach direction of a mid lane
    approaching: bool
    lane_id: int
    length: float


class ItscpGrid:
    """Lane graph of `_make_road` (_env.py:225-439): per intersection 4 corners x (approaching, leaving) x num_lane
    access lanes, then the connecting lanes (straight for every approaching lane, right turn for the outermost
    one; left turns are disabled in the reference), then the links between neighbouring intersections."""

    def __init__(self, num_intersection: int = 1, num_lane: int = 3, lane_length: float = 20.0, cell_length: float = 5.0):
        n, nl = int(num_intersection), int(num_lane)
        self.num_intersection, self.num_lane = n, nl
        self.lane_length, self.cell_length = float(lane_length), float(cell_length)
        outer = (LANE_WIDTH + 10.0) + LANE_WIDTH * (nl - 3 + 0.5)
        access = float(lane_length)
        lanes: List[LaneInfo] = []
        links: List[Tuple[int, int]] = []
        index = {}
        ends = {}                         # lane -> (start xy, end xy) of the drawn segment
        loc_of = {(0, True): "south", (0, False): "east", (1, True): "west", (1, False): "south",
                  (2, True): "north", (2, False): "west", (3, True): "east", (3, False): "north"}
        straight = {"north": "south", "west": "east", "east": "west", "south": "north"}
        right = {"north": "west", "west": "south", "east": "north", "south": "east"}

        def add(info, a, b):
            info.length = float(np.linalg.norm(np.asarray(b) - np.asarray(a)))
            index[(info.row, info.col, info.loc, info.ploc, info.approaching, info.lane_id)] = len(lanes)
            ends[len(lanes)] = (np.asarray(a, dtype=float), np.asarray(b, dtype=float))
            lanes.append(info)
            return len(lanes) - 1

        for row in range(n):
            for col in range(n):
                center = np.array([col * (outer + access), row * (outer + access)]) * 2.0
                approaching_ids = []
                for corner in range(4):
                    ang = math.radians(90 * corner)
                    rot = np.array([[math.cos(ang), -math.sin(ang)], [math.sin(ang), math.cos(ang)]])
                    for app in (True, False):
                        for k in range(nl):
                            a = np.array([LANE_WIDTH * (k + 0.5), access + outer])
                            b = np.array([LANE_WIDTH * (k + 0.5), outer])
                            if not app:
                                a, b = a[::-1], b[::-1]
                            i = add(LaneInfo(row, col, loc_of[(corner, app)], None, app, k, 0.0), center + rot @ a, center + rot @ b)
                            if app:
                                approaching_ids.append(i)
                idx = 0
                for i in approaching_ids:
                    li = lanes[i]
                    for turn in ("straight", "right"):
                        if turn == "right" and li.lane_id != nl - 1:
                            continue
                        nloc = straight[li.loc] if turn == "straight" else right[li.loc]
                        j = index[(row, col, nloc, None, False, li.lane_id)]
                        m = add(LaneInfo(row, col, "mid", li.loc, True, idx, 0.0), ends[i][1], ends[j][1])
                        idx += 1
                        links.append((i, m)); links.append((m, j))
        for row in range(n):
            for col in range(n):
                for cond, here, there, dr, dc in ((row > 0, "north", "south", -1, 0), (col > 0, "west", "east", 0, -1)):
                    if not cond:
                        continue
                    for app in (True, False):
                        for k in range(nl):
                            cur = index[(row, col, here, None, app, k)]
                            con = index[(row + dr, col + dc, there, None, not app, k)]
                            links.append((con, cur) if app else (cur, con))
        self.lanes, self.links = lanes, links
        self._sig_tables = {}
        self.L = len(lanes)
        self.num_cell = [max(1, math.ceil(l.length / self.cell_length)) for l in lanes]     # MacroLane.__init__, _macro_lane.py:38-44
        self.dx = [l.length / c for l, c in zip(lanes, self.num_cell)]

    def topology(self, device) -> MacroNetTopology:
        return MacroNetTopology(self.num_cell, self.dx, self.links, device, MODE_ITSCP)

    def boundary_lanes(self) -> List[int]:
        """Lanes without predecessor (they receive the scheduled inflow)."""
        has_prev = {b for _, b in self.links}
        return [i for i in range(self.L) if i not in has_prev]

    # ------------------------------------------------------------------ signals
    def signals(self, action: torch.Tensor, num_frames: int, frames_per_signal: int, soft: bool = True) -> torch.Tensor:
        """lane_signal of every lane at every frame: [R, T, L] from action [R, n_phase * n^2] (lane_signal_info's
        next_signal, _env.py:885-962; ``simulator.lane_signal[id] = signal_info[1]``, :600-603)."""
        R, A = action.shape
        n2 = self.num_intersection ** 2
        n_phase = A // n2
        dev = action.device
        key = (str(dev), action.dtype)
        tab = self._sig_tables.get(key)
        if tab is None:      # approaching access lanes: their lane index, intersection cell and sign (+1 west / east, -1 north / south)
            lanes = [(l, i.row * self.num_intersection + i.col, 1.0 if i.loc in ("west", "east") else -1.0)
                     for l, i in enumerate(self.lanes) if i.loc != "mid" and i.approaching]
            tab = (torch.tensor([x[0] for x in lanes], dtype=torch.long, device=dev),
                   torch.tensor([x[1] for x in lanes], dtype=torch.long, device=dev),
                   torch.tensor([x[2] for x in lanes], dtype=action.dtype, device=dev))
            self._sig_tables[key] = tab
        lane_idx, cell_idx, sign = tab
        t = torch.arange(num_frames, device=dev)
        phase = torch.clamp(t // frames_per_signal, max=n_phase - 1)
        progress = torch.clamp((t % frames_per_signal).to(action.dtype) / frames_per_signal, max=1.0)
        sig = torch.ones((R, num_frames, self.L), dtype=action.dtype, device=dev)
        if lane_idx.numel():
            col = phase[:, None] * n2 + cell_idx[None, :]                        # [T, lanes] action column of every (frame, lane)
            d = (action[:, col] - progress[None, :, None]) * sign                # a - progress (W/E) or progress - a (N/S)
            if soft:
                s = torch.sigmoid(torch.clamp(d * 32.0, -16.0, 16.0))            # dmath.sigmoid, operation.py:3-30
            else:
                s = (d > 0).to(action.dtype)
            sig = sig.index_copy(2, lane_idx, s)
        return sig


# ---------------------------------------------------------------------- queue-length reward, exact constants
def queue_constants(u_states: torch.Tensor, static_speed: float, window: int = 100_000) -> torch.Tensor:
    """Sigmoid constant 16 / |running mean| the reference applies to every cell sample (_env.py:557-575):
    before each sigmoid the sample (static_speed - u) is appended to a RunningMean over the last `window` samples
    (common/rms.py), visiting frames, lanes and cells in order.  u_states [T, NC] (one replica) or [R, T, NC] (one
    RunningMean per replica), states AFTER each step, cells lane by lane -> constants of the same shape.  The reference
    accumulates in float32; this is float64."""
    shape = u_states.shape
    d = (static_speed - u_states.detach()).to(torch.float64)
    d = d.reshape(1, -1) if d.dim() == 2 else d.reshape(shape[0], -1)
    cs = torch.cumsum(d, 1)
    n = torch.arange(1, d.shape[1] + 1, device=d.device)
    lo = n - window
    drop = torch.where(lo > 0, cs[:, torch.clamp(lo - 1, min=0)], torch.zeros_like(cs))
    cnt = torch.clamp(n, max=window).to(torch.float64)
    mean = (cs - drop) / cnt
    return (16.0 / mean.abs()).reshape(shape).to(u_states.dtype)


def queue_reward(states: torch.Tensor, topo: MacroNetTopology, dt: float, veh_len: float, static_speed: float,
                 constants: torch.Tensor) -> torch.Tensor:
    """-sum_t sum_lanes (sum_cells sigmoid(k (static - u)) r dx / len)^2 dt (_env.py:618-648,770-797) of the states
    AFTER each step.  states [T+1, R, 3, NC]; constants [R, T, NC] (or broadcastable).  Returns [R]."""
    r = states[1:, :, 0].transpose(0, 1)         # [R, T, NC]
    u = states[1:, :, 2].transpose(0, 1)
    z = torch.clamp((static_speed - u) * constants, -16.0, 16.0)
    w = topo.dx(states.dtype)[topo.lane_of_cell()] / veh_len
    per_cell = torch.sigmoid(z) * (r * w)
    q = torch.zeros(per_cell.shape[:2] + (topo.L,), dtype=states.dtype, device=states.device)
    q = q.index_add(2, topo.lane_of_cell(), per_cell)
    return -(q ** 2.0).sum(dim=(1, 2)) * dt


class ItscpBatch:
    """R candidate signal plans / inflow schedules of one ITSCP grid, evaluated and differentiated together.

    ``rollout(action, incoming, routes)`` returns (reward [R], states); reward is the reference's queue-length
    reward.  ``exact_constants=True`` reproduces the running-mean sigmoid constants sample by sample (reward and
    its adjoint are then computed by torch ops over the stored states, injected through ``g_states``);
    ``False`` fuses the reward into the kernels with one constant per frame (``qk``), the fast path."""

    def __init__(self, grid: ItscpGrid, device, speed_limit: float = 60.0, simulation_frequency: int = 30,
                 signal_length: float = 2.0, vehicle_length: float = 5.0, static_speed: float = 0.2, dtype=torch.float64):
        self.grid, self.device, self.dtype = grid, torch.device(device), dtype
        self.topo = grid.topology(device)
        self.umax, self.freq = float(speed_limit), int(simulation_frequency)
        self.frames_per_signal = int(simulation_frequency * signal_length)
        self.dt = 1.0 / simulation_frequency
        self.veh_len, self.static_speed = float(vehicle_length), float(static_speed)

    def rollout(self, action, incoming, routes, num_frames: int, differentiable: bool = True, r0=None, u0=None,
                exact_constants: bool = True, qk=None, flags: Optional[_lib.Flags] = None):
        R = action.shape[0]
        sig = self.grid.signals(action, num_frames, self.frames_per_signal, soft=differentiable)
        NC = self.topo.NC
        if r0 is None:      # lanes start empty: ARZ.FullQ(u_max), _arz.py:59-63
            r0 = torch.zeros((R, NC), dtype=self.dtype, device=self.device)
            u0 = torch.full((R, NC), self.umax, dtype=self.dtype, device=self.device)
        flags = flags or _lib.Flags(self.device)
        states, reward = net_rollout(self.topo, r0, u0, self.umax, self.dt, num_frames, sig=sig, incoming=incoming,
                                     route=routes, soft=differentiable, qk=None if exact_constants else qk,
                                     veh_len=self.veh_len, static_speed=self.static_speed, flags=flags)
        if exact_constants:
            u_after = states[1:, :, 2].transpose(0, 1)
            if differentiable:
                k = queue_constants(u_after, self.static_speed)
            else:   # hard test: speed < static_speed (_env.py:576-586)
                k = torch.full_like(u_after, 1e30)
            reward = queue_reward(states, self.topo, self.dt, self.veh_len, self.static_speed, k)
        return reward, states
