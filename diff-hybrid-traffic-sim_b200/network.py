"""Connected macro networks, batched over replicas: host side of ``dhts_net_rollout_{fwd,bwd}_*``.

One launch steps R replicas (scenarios, candidate signal plans, CMA-ES populations) of one network of
macro lanes for T steps; the adjoint launch returns gradients wrt the initial state, the lane signals
and the boundary inflow.  It stands for T x ``RoadNetwork.forward`` (road/network/road_network.py:79-111)
with ``get_macro_boundary`` (:299-362) -- and, in ITSCP mode, ``ItscpRoadNetwork.setup_macro_boundary``
(example/control/itscp/_simulator.py:56-142) -- plus the autograd chain through the per-lane operators.

``MacroNetTopology`` is the lane graph in the kernel's layout; build it from explicit lists or from any
object graph with the reference's lane surface (``RoadNetwork.lane`` dict of macro lanes with
``prev_lane`` / ``next_lane`` dicts, ``num_cell`` and ``cell_length``): ``MacroNetTopology.from_network``.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, creal, ptr, stream_ptr, suffix
from .ops import EPSILON

MODE_PLAIN, MODE_ITSCP = 0, 1


class _TopoStruct(ctypes.Structure):
    """``dhts_net_topology`` of include/dhts.h."""
    _fields_ = [("L", ctypes.c_int), ("NC", ctypes.c_int), ("n_own", ctypes.c_int),
                ("cell_off", ctypes.c_void_p), ("nadj", ctypes.c_void_p), ("one_adj", ctypes.c_void_p),
                ("adj_off", ctypes.c_void_p), ("adj", ctypes.c_void_p), ("own_slot", ctypes.c_void_p)]


class MacroNetTopology:
    """Lane graph of a macro network: cells per lane, cell length per lane, directed lane links.

    ``own`` selects which (lane, side) pairs keep their own ghost record: every side without a neighbour
    does (road_network.py:312-321); ``mode=MODE_ITSCP`` drops the left side of lanes without predecessor,
    whose ghost is the scheduled inflow instead (_simulator.py:68-73).
    """

    def __init__(self, num_cell: Sequence[int], cell_length: Sequence[float], links: Sequence[Tuple[int, int]],
                 device, mode: int = MODE_PLAIN):
        L = len(num_cell)
        assert L >= 1 and len(cell_length) == L and all(n >= 1 for n in num_cell)
        self.L, self.mode = L, int(mode)
        self.num_cell = [int(n) for n in num_cell]
        self.cell_length = [float(x) for x in cell_length]
        self.links = [(int(a), int(b)) for a, b in links]
        cell_off = [0]
        for n in self.num_cell:
            cell_off.append(cell_off[-1] + n)
        self.NC = cell_off[-1]
        prev: List[List[int]] = [[] for _ in range(L)]
        nxt: List[List[int]] = [[] for _ in range(L)]
        for a, b in self.links:           # insertion order = the reference's dict order (connect_lane, road_network.py:66-77)
            assert 0 <= a < L and 0 <= b < L
            if b not in nxt[a]:
                nxt[a].append(b)
            if a not in prev[b]:
                prev[b].append(a)
        self.prev, self.next = prev, nxt
        nadj = [len(p) for p in prev] + [len(n) for n in nxt]
        one = [p[0] if len(p) == 1 else -1 for p in prev] + [n[0] if len(n) == 1 else -1 for n in nxt]
        adj, adj_off = [], []
        for lists in (prev, nxt):
            for li in lists:
                adj_off.append(len(adj)); adj.extend(li)
            adj_off.append(len(adj))
        own_slot, n_own = [], 0
        for side, lists in enumerate((prev, nxt)):
            for l in range(L):
                keeps = len(lists[l]) == 0 and not (self.mode == MODE_ITSCP and side == 0)
                own_slot.append(n_own if keeps else -1)
                n_own += int(keeps)
        self.n_own = n_own
        self.own_slot = own_slot
        self.cell_off = cell_off
        self.device = torch.device(device)
        parts = {"cell_off": cell_off, "nadj": nadj, "one_adj": one, "adj_off": adj_off, "adj": adj or [0],
                 "own_slot": own_slot}
        self._dev: Dict[str, torch.Tensor] = {}
        if self.device.type == "cuda":
            for k, v in parts.items():
                self._dev[k] = torch.tensor(v, dtype=torch.int32, device=self.device)
            self._struct = _TopoStruct(L, self.NC, n_own, *(self._dev[k].data_ptr() for k in
                                                             ("cell_off", "nadj", "one_adj", "adj_off", "adj", "own_slot")))
        self.host = parts
        self._dx: Dict[torch.dtype, torch.Tensor] = {}

    # ------------------------------------------------------------------ builders
    @classmethod
    def from_network(cls, network, device, mode: int = MODE_PLAIN) -> "MacroNetTopology":
        """Lane graph of a RoadNetwork-like object whose lanes are all macro lanes (ids 0..L-1)."""
        lanes = network.lane
        ids = sorted(lanes.keys())
        assert ids == list(range(len(ids))), "lane ids must be 0..L-1 (RoadNetwork.add_lane numbers them so)"
        assert all(lanes[i].is_macro() for i in ids), "the fused network rollout steps macro lanes only"
        links = [(i, j) for i in ids for j in lanes[i].next_lane.keys()]
        topo = cls([lanes[i].num_cell for i in ids], [lanes[i].cell_length for i in ids], links, device, mode)
        return topo

    def struct_ptr(self):
        if self.device.type != "cuda":
            raise RuntimeError("the network rollout runs on CUDA only (no CPU fallback)")
        return ctypes.byref(self._struct)

    def dx(self, dtype) -> torch.Tensor:
        if dtype not in self._dx:
            self._dx[dtype] = torch.tensor(self.cell_length, dtype=dtype, device=self.device)
        return self._dx[dtype]

    def lane_of_cell(self) -> torch.Tensor:
        return torch.repeat_interleave(torch.arange(self.L), torch.tensor(self.num_cell)).to(self.device)

    def route_table(self, routes) -> torch.Tensor:
        """[T][2][L] int32 (prev lane, next lane) from a list of MacroRoute-like objects (one per step) exposing
        ``get_prev_lane`` / ``get_next_lane`` (road/network/route.py:18-38)."""
        rows = [[[r.get_prev_lane(l) for l in range(self.L)], [r.get_next_lane(l) for l in range(self.L)]] for r in routes]
        return torch.tensor(rows, dtype=torch.int32, device=self.device)

    def default_own(self, R: int, dtype, umax: float) -> torch.Tensor:
        """Initial own ghost records (r, u) = (0, u_max): ``ARZ.FullQ(u_max)`` (model/macro/_arz.py:59-63)."""
        o = torch.zeros((R, max(self.n_own, 0), 2), dtype=dtype, device=self.device)
        o[..., 1] = umax
        return o


class NetRolloutFn(torch.autograd.Function):
    """(r0, y0, u0, own0, sig, incoming) -> (hist [T+1,R,3,NC], reward [R])."""

    @staticmethod
    def forward(ctx, r0, y0, u0, own0, sig, incoming, topo: MacroNetTopology, route, qk, umax, dt, veh_len,
                static_speed, steps, soft, flags):
        dev = _lib.require_cuda(r0, y0, u0, own0, sig, incoming, route, qk, flags)
        ctx.set_materialize_grads(False)      # an output nobody differentiates stays None in backward (no zero-filled history)
        c = lambda t: None if t is None else t.contiguous()
        r0, y0, u0, own0, sig, incoming, route, qk = map(c, (r0, y0, u0, own0, sig, incoming, route, qk))
        R, NC = r0.shape
        assert NC == topo.NC, "state rows must hold the network's cells lane by lane"
        dtype = r0.dtype
        steps = int(steps)
        L = topo.L
        if topo.mode == MODE_ITSCP:
            assert sig is not None and incoming is not None and sig.shape == (R, steps, L) == incoming.shape
        per_rep = 0
        if route is not None:
            assert route.dtype == torch.int32 and route.shape[-3:] == (steps, 2, L)
            per_rep = int(route.dim() == 4 and route.shape[0] == R and R > 1)
        if qk is not None:
            assert qk.shape == (steps,) and qk.dtype == dtype
        if own0 is None:
            own0 = topo.default_own(R, dtype, umax)
        hist = torch.empty((steps + 1, R, 3, NC), dtype=dtype, device=dev)
        ownh = torch.empty((steps + 1, R, max(topo.n_own, 1), 2), dtype=dtype, device=dev)
        if topo.n_own == 0:
            ownh.zero_()
        reward = torch.zeros((R,), dtype=dtype, device=dev)
        fn = getattr(_lib.load(), "dhts_net_rollout_fwd_" + suffix(dtype))
        with torch.cuda.device(dev):
            check(fn(topo.struct_ptr(), ptr(topo.dx(dtype)), ptr(route), per_rep, ptr(sig), ptr(incoming), ptr(qk),
                     creal(dtype, umax), creal(dtype, dt), creal(dtype, veh_len), creal(dtype, static_speed), steps, R,
                     topo.mode, int(bool(soft)), ptr(r0), ptr(y0), ptr(u0), ptr(own0 if topo.n_own else None),
                     ptr(hist), ptr(ownh if topo.n_own else None), ptr(reward if qk is not None else None), ptr(flags),
                     stream_ptr(dev)), "dhts_net_rollout_fwd")
        ctx.save_for_backward(hist, ownh, sig, incoming, route, qk)
        ctx.cfg = (topo, per_rep, float(umax), float(dt), float(veh_len), float(static_speed), steps, int(bool(soft)), R)
        ctx.flags = flags
        return hist, reward

    @staticmethod
    def backward(ctx, g_hist, g_reward):
        hist, ownh, sig, incoming, route, qk = ctx.saved_tensors
        topo, per_rep, umax, dt, veh_len, static_speed, steps, soft, R = ctx.cfg
        dev, dtype = hist.device, hist.dtype
        NC, L = topo.NC, topo.L
        g_states = None
        if g_hist is not None and steps > 0:
            g_states = g_hist[1:].contiguous()
        g_rew = None if (g_reward is None or qk is None) else g_reward.contiguous()
        g_r0 = torch.empty((R, NC), dtype=dtype, device=dev); g_y0 = torch.empty_like(g_r0); g_u0 = torch.empty_like(g_r0)
        g_own0 = torch.zeros((R, max(topo.n_own, 1), 2), dtype=dtype, device=dev)
        itscp = topo.mode == MODE_ITSCP
        g_sig = torch.zeros((R, steps, L), dtype=dtype, device=dev) if itscp else None
        g_inc = torch.zeros((R, steps, L), dtype=dtype, device=dev) if itscp else None
        fn = getattr(_lib.load(), "dhts_net_rollout_bwd_" + suffix(dtype))
        with torch.cuda.device(dev):
            check(fn(topo.struct_ptr(), ptr(topo.dx(dtype)), ptr(route), per_rep, ptr(sig), ptr(incoming), ptr(qk),
                     creal(dtype, umax), creal(dtype, dt), creal(dtype, veh_len), creal(dtype, static_speed), steps, R,
                     topo.mode, soft, ptr(hist), ptr(ownh if topo.n_own else None), ptr(g_states), ptr(g_rew), ptr(g_r0),
                     ptr(g_y0), ptr(g_u0), ptr(g_own0 if topo.n_own else None), ptr(g_sig), ptr(g_inc), ptr(ctx.flags),
                     stream_ptr(dev)), "dhts_net_rollout_bwd")
        if g_hist is not None:
            g_r0 = g_r0 + g_hist[0, :, 0]; g_y0 = g_y0 + g_hist[0, :, 1]; g_u0 = g_u0 + g_hist[0, :, 2]
        need = ctx.needs_input_grad
        out = (g_r0, g_y0, g_u0, g_own0[:, :topo.n_own] if topo.n_own else None, g_sig, g_inc)
        return tuple(g if need[i] else None for i, g in enumerate(out)) + (None,) * 10


def net_rollout(topo: MacroNetTopology, r0, u0, umax: float, dt: float, steps: int, *, sig=None, incoming=None,
                route=None, own0=None, soft: bool = True, qk=None, veh_len: float = 5.0, static_speed: float = 0.2,
                flags: Optional[_lib.Flags] = None):
    """`steps` x RoadNetwork.forward over R replicas of a connected macro network.

    r0, u0 [R, NC] initial density / speed of every cell (``set_state_vector_u`` per lane, lane by lane);
    sig, incoming [R, steps, L] (ITSCP mode); route [steps, 2, L] or [R, steps, 2, L] int32 (``topo.route_table``);
    own0 [R, n_own, 2] initial own ghost records (default (0, u_max)); qk [steps]: enables the fused queue reward.
    Returns (states [steps+1, R, 3, NC] holding (r, y, u) before each step and after the last, reward [R])."""
    flags = flags or _lib.Flags(r0.device)
    ueq = umax * (1.0 - torch.sqrt(torch.clamp(r0, min=0.0) + EPSILON))
    y0 = r0 * (u0 - ueq)                           # set_r_u, _arz.py:82-86 (autograd: true derivative)
    return NetRolloutFn.apply(r0, y0, u0, own0, sig, incoming, topo, route, qk, float(umax), float(dt), float(veh_len),
                              float(static_speed), int(steps), bool(soft), flags.t)
