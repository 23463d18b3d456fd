"""Connected HYBRID networks (macro ARZ lanes + micro IDM lanes), batched over replicas: host side of
``dhts_hyb_rollout_{fwd,bwd}_*``.

One launch steps R replicas of one network for T steps, one more returns the gradients.  It stands for
T x ``RoadNetwork.forward`` (road/network/road_network.py:79-173) on a network that mixes ``dMacroLane`` and
``dMicroLane`` objects -- ghost cells from neighbouring macro lanes, head-vehicle leaders along vehicle routes
(:429-580), the ITSCP signal blends (example/control/itscp/_simulator.py:56-276) and every ``Conversion.*``
(road/network/conversion.py) in lane-id order -- plus the autograd chain through the per-lane operators.

What stays an INPUT: the reference draws a random route for every spawned vehicle with ``np.random``
(``create_random_route``, road_network.py:604-646).  Here ``spawn_route[m][k]`` names the route of the k-th vehicle
spawned into micro lane m; ``HybridNetTopology.random_spawn_routes`` draws them the same way (uniform next lane at
every hop).  Only the part of a route up to its first macro lane matters (the vehicle is absorbed there), so routes
are stored as "micro prefix + first macro lane".

IDM parameters are per VEHICLE: a call takes a table of parameter sets (``veh_params`` [NP, 6]) and every initial vehicle
names its set (``make_aux0(pid0=...)``; MicroVehicle's own attributes, road/vehicle/micro_vehicle.py:74-122); vehicles the
network creates itself (macro->micro spawns, waiting-list entries) are ``default_micro_vehicle``s = set 0, as in the
reference (conversion.py:53-57, _env.py:202-219).

ITSCP MICRO mode (every lane a plain ``MicroLane``, _env.py:484-488): ``sources=True`` marks the micro lanes without
predecessor as fed from a waiting list (_simulator.py:153-174).  The uniform draws the reference takes from
``np.random`` while a lane has room are an INPUT (``src_rand``, consumption order), like the waiting routes
(``spawn_route`` rows in pop order).

Restriction: with ``enumerate_routes=True`` (default) a vehicle route on a cycle of micro lanes ends where every
successor has been visited (``cyclic_micro``; the reference would let it circulate for up to 32 hops);
``enumerate_routes=False`` registers routes as they are named (``route_id``) and takes them as they are.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, creal, ptr, stream_ptr, suffix
from .network import MODE_ITSCP, MODE_PLAIN
from .ops import EPSILON

ROUTE_LEN = 32          # MAX_ROUTE_LENGTH, road_network.py:15


class _HybTopoStruct(ctypes.Structure):
    """``dhts_hyb_topology`` of include/dhts.h."""
    _fields_ = [("L", ctypes.c_int), ("NC", ctypes.c_int), ("n_own", ctypes.c_int),
                ("cell_off", ctypes.c_void_p), ("nadj", ctypes.c_void_p), ("one_adj", ctypes.c_void_p),
                ("adj_off", ctypes.c_void_p), ("adj", ctypes.c_void_p), ("own_slot", ctypes.c_void_p),
                ("ML", ctypes.c_int), ("cap", ctypes.c_int), ("NCAP", ctypes.c_int), ("NG", ctypes.c_int),
                ("NGL", ctypes.c_int), ("NR", ctypes.c_int), ("RLEN", ctypes.c_int), ("MAXT", ctypes.c_int),
                ("kind", ctypes.c_void_p), ("mic_of", ctypes.c_void_p), ("mic_lane", ctypes.c_void_p),
                ("cap_off", ctypes.c_void_p), ("cap_lane", ctypes.c_void_p), ("grp_off", ctypes.c_void_p),
                ("grp_lane", ctypes.c_void_p), ("routes", ctypes.c_void_p), ("src", ctypes.c_void_p)]


def default_vehicle_params(speed_limit: float, length: float = 5.0) -> List[float]:
    """(a_max, a_pref, v_target, s0, T, length) of ``MicroVehicle.default_micro_vehicle`` (micro_vehicle.py:30-72)."""
    return [speed_limit * 1.0, speed_limit * 0.8, speed_limit * 0.9, length * 0.1, 0.1, length]


class HybridNetTopology:
    """Lane graph of a hybrid network.  ``kind[l]`` 0 = macro, 1 = micro; ``num_cell`` / ``cell_length`` are read for
    macro lanes only; ``links`` in the reference's ``connect_lane`` order."""

    def __init__(self, kind: Sequence[int], num_cell: Sequence[int], cell_length: Sequence[float],
                 lane_length: Sequence[float], links: Sequence[Tuple[int, int]], device, mode: int = MODE_PLAIN,
                 veh_cap: int = 8, veh_len: float = 5.0, max_routes: int = 4096, sources: bool = False,
                 enumerate_routes: bool = True):
        L = len(kind)
        assert L >= 1 and len(num_cell) == L == len(cell_length) == len(lane_length)
        self.L, self.mode = L, int(mode)
        self.kind = [int(k) for k in kind]
        self.num_cell = [0 if self.kind[l] else int(num_cell[l]) for l in range(L)]
        assert all(self.kind[l] or self.num_cell[l] >= 1 for l in range(L))
        self.cell_length = [1.0 if self.kind[l] else float(cell_length[l]) for l in range(L)]
        self.lane_length = [float(x) for x in lane_length]
        self.links = [(int(a), int(b)) for a, b in links]
        self.veh_cap, self.veh_len = int(veh_cap), float(veh_len)
        cell_off = [0]
        for n in self.num_cell:
            cell_off.append(cell_off[-1] + n)
        self.NC = cell_off[-1]
        self.cell_off = cell_off
        prev: List[List[int]] = [[] for _ in range(L)]
        nxt: List[List[int]] = [[] for _ in range(L)]
        for a, b in self.links:
            assert 0 <= a < L and 0 <= b < L
            if b not in nxt[a]:
                nxt[a].append(b)
            if a not in prev[b]:
                prev[b].append(a)
        self.prev, self.next = prev, nxt
        self.micro = [l for l in range(L) if self.kind[l]]
        self.ML = len(self.micro)
        mic_of = [-1] * L
        for i, l in enumerate(self.micro):
            mic_of[l] = i
        self.mic_of = mic_of
        self.sources = bool(sources)
        if self.mode == MODE_ITSCP and not self.sources:
            assert all(prev[l] for l in self.micro), "ITSCP mode: micro lanes without predecessor are waiting-list sources (sources=True)"
        assert not self.sources or self.mode == MODE_ITSCP, "waiting-list sources belong to the ITSCP simulator"
        self.src = [int(self.sources and not prev[l]) for l in self.micro]
        nadj = [len(p) for p in prev] + [len(n) for n in nxt]
        one = [p[0] if len(p) == 1 else -1 for p in prev] + [n[0] if len(n) == 1 else -1 for n in nxt]
        adj, adj_off = [], []
        for lists in (prev, nxt):
            for li in lists:
                adj_off.append(len(adj)); adj.extend(li)
            adj_off.append(len(adj))
        # own ghost records: sides without neighbour (road_network.py:312-321) and sides that can face a micro lane (:353-362)
        own_slot, n_own = [], 0
        for side, lists in enumerate((prev, nxt)):
            for l in range(L):
                keeps = (not self.kind[l]) and ((len(lists[l]) == 0 and not (self.mode == MODE_ITSCP and side == 0))
                                                or any(self.kind[x] for x in lists[l]))
                own_slot.append(n_own if keeps else -1)
                n_own += int(keeps)
        self.n_own, self.own_slot = n_own, own_slot
        # flux capacitors: MacroLane.flux_capacitor[next micro lane id] (_macro_lane.py:215-225)
        cap_off, cap_lane = [0], []
        for l in range(L):
            if not self.kind[l]:
                cap_lane.extend(x for x in nxt[l] if self.kind[x])
            cap_off.append(len(cap_lane))
        self.cap_off, self.cap_lane, self.NCAP = cap_off, cap_lane, len(cap_lane)
        # conversion groups: lanes whose conversions can touch the same lane within a step end up in one group
        root = list(range(L))

        def find(i):
            while root[i] != i:
                root[i] = root[root[i]]; i = root[i]
            return i

        for l in range(L):
            for x in nxt[l]:
                if self.kind[l] or self.kind[x]:
                    root[find(l)] = find(x)
        conv = [l for l in range(L) if self.kind[l] or cap_off[l + 1] > cap_off[l]]
        groups: Dict[int, List[int]] = {}
        for l in conv:
            groups.setdefault(find(l), []).append(l)
        self.groups = sorted((sorted(g) for g in groups.values()), key=lambda g: g[0])
        grp_off, grp_lane = [0], []
        for g in self.groups:
            grp_lane.extend(g); grp_off.append(len(grp_lane))
        # Cycles in the micro sub-graph (e.g. the four centre intersections of a 4 x 4 ITSCP grid): once every successor
        # is on the route already, the reference's create_random_route keeps its first choice and lets the vehicle
        # circulate for up to 32 hops (road_network.py:628-640).  Those revisiting routes grow combinatorially and are
        # NOT enumerated here: a route ends where every successor has been visited, and the vehicle leaves there through
        # micro_to_none.  `cyclic_micro` records that this network is affected (the drop-in RoadNetwork, which draws
        # routes exactly like the reference, is the path to use when circulating vehicles matter).
        state = {}

        def cyclic(l):
            if state.get(l) == 1:
                return True
            if state.get(l) == 2:
                return False
            state[l] = 1
            hit = any(self.kind[x] and cyclic(x) for x in nxt[l])
            state[l] = 2
            return hit

        self.cyclic_micro = any(cyclic(l) for l in self.micro)
        # vehicle routes: micro prefix + first macro lane, enumerated from every micro lane
        self.routes: List[Tuple[int, ...]] = []
        self.route_index: Dict[Tuple[int, ...], int] = {}

        def walk(path):
            l = path[-1]
            if not self.kind[l] or not nxt[l] or len(path) >= ROUTE_LEN:
                self.route_index[tuple(path)] = len(self.routes); self.routes.append(tuple(path))
                assert len(self.routes) <= max_routes, "too many distinct vehicle routes: raise max_routes"
                return
            for x in nxt[l]:
                if x in path:        # create_random_route avoids revisits when it can (road_network.py:628-640)
                    continue
                walk(path + [x])
            if all(x in path for x in nxt[l]):
                self.route_index[tuple(path)] = len(self.routes); self.routes.append(tuple(path))

        self.enumerate_routes = bool(enumerate_routes)
        if self.enumerate_routes:
            for l in self.micro:
                walk([l])
        route_rows = [list(r) + [-1] * (ROUTE_LEN - len(r)) for r in self.routes] or [[-1] * ROUTE_LEN]
        dxs = [self.cell_length[l] for l in range(L) if not self.kind[l]]
        self.MAXT = int(math.ceil(self.veh_len / min(dxs))) + 1 if dxs else 1
        self.device = torch.device(device)
        parts = {"cell_off": cell_off, "nadj": nadj, "one_adj": one, "adj_off": adj_off, "adj": adj or [0], "own_slot": own_slot,
                 "kind": self.kind, "mic_of": mic_of, "mic_lane": self.micro or [0], "cap_off": cap_off,
                 "cap_lane": cap_lane or [0], "grp_off": grp_off, "grp_lane": grp_lane or [0],
                 "routes": [x for row in route_rows for x in row], "src": self.src or [0]}
        self.host = parts
        self._dev: Dict[str, torch.Tensor] = {}
        self._real: Dict[Tuple[str, torch.dtype], torch.Tensor] = {}
        if self.device.type == "cuda":
            for k, v in parts.items():
                self._dev[k] = torch.tensor(v, dtype=torch.int32, device=self.device)
            d = self._dev
            self._struct = _HybTopoStruct(
                L, self.NC, n_own, d["cell_off"].data_ptr(), d["nadj"].data_ptr(), d["one_adj"].data_ptr(), d["adj_off"].data_ptr(),
                d["adj"].data_ptr(), d["own_slot"].data_ptr(), self.ML, self.veh_cap, self.NCAP, len(self.groups), len(grp_lane),
                len(self.routes), ROUTE_LEN, self.MAXT, d["kind"].data_ptr(), d["mic_of"].data_ptr(), d["mic_lane"].data_ptr(),
                d["cap_off"].data_ptr(), d["cap_lane"].data_ptr(), d["grp_off"].data_ptr(), d["grp_lane"].data_ptr(),
                d["routes"].data_ptr(), d["src"].data_ptr() if self.sources else None)
        s = self.ML * self.veh_cap
        self.A_P, self.A_V, self.A_A, self.A_RID, self.A_CUR, self.A_PID = 0, s, 2 * s, 3 * s, 4 * s, 5 * s
        self.A_FRONT = 6 * s; self.A_CNT = self.A_FRONT + self.ML; self.A_NSP = self.A_CNT + self.ML
        self.A_CAP = self.A_NSP + self.ML; self.A_RMS = self.A_CAP + self.NCAP; self.A_DRAW = self.A_RMS + 2
        self.AUX = self.A_DRAW + 1

    # ------------------------------------------------------------------ builders
    @classmethod
    def from_network(cls, network, device, mode: int = MODE_PLAIN, **kw) -> "HybridNetTopology":
        """Lane graph of a RoadNetwork-like object (lane ids 0..L-1, ``is_macro()``, ``num_cell``, ``cell_length``,
        ``length``, ``next_lane``)."""
        lanes = network.lane
        ids = sorted(lanes.keys())
        assert ids == list(range(len(ids))), "lane ids must be 0..L-1 (RoadNetwork.add_lane numbers them so)"
        kind = [0 if lanes[i].is_macro() else 1 for i in ids]
        links = [(i, j) for i in ids for j in lanes[i].next_lane.keys()]
        return cls(kind, [lanes[i].num_cell if not kind[i] else 0 for i in ids],
                   [lanes[i].cell_length if not kind[i] else 1.0 for i in ids], [lanes[i].length for i in ids], links, device,
                   mode, veh_len=getattr(network, "vehicle_length", 5.0), **kw)

    def struct_ptr(self):
        if self.device.type != "cuda":
            raise RuntimeError("the network rollout runs on CUDA only (no CPU fallback)")
        self._sync_routes()
        return ctypes.byref(self._struct)

    def real(self, name: str, dtype) -> torch.Tensor:
        key = (name, dtype)
        if key not in self._real:
            src = {"dx": self.cell_length, "lane_len": self.lane_length}[name]
            self._real[key] = torch.tensor(src, dtype=dtype, device=self.device)
        return self._real[key]

    def lane_of_cell(self) -> torch.Tensor:
        return torch.repeat_interleave(torch.arange(self.L), torch.tensor(self.num_cell)).to(self.device)

    def route_table(self, routes) -> torch.Tensor:
        """[T][2][L] int32 (prev lane, next lane) from MacroRoute-like objects, one per step (road/network/route.py:18-38)."""
        rows = [[[r.get_prev_lane(l) for l in range(self.L)], [r.get_next_lane(l) for l in range(self.L)]] for r in routes]
        return torch.tensor(rows, dtype=torch.int32, device=self.device)

    def route_id(self, path: Sequence[int]) -> int:
        """Id of a vehicle route given as the reference's lane list: cut after the first macro lane.  With
        ``enumerate_routes=False`` an unknown route is registered (the device table is re-uploaded before the next launch)."""
        cut = []
        for l in path:
            if l < 0:
                break
            cut.append(int(l))
            if not self.kind[l]:
                break
        key = tuple(cut)
        if key not in self.route_index:
            if self.enumerate_routes:
                raise KeyError(key)
            assert 1 <= len(key) <= ROUTE_LEN and self.kind[key[0]] == 1
            self.route_index[key] = len(self.routes); self.routes.append(key)
            self._routes_dirty = True
        return self.route_index[key]

    def _sync_routes(self):
        if getattr(self, "_routes_dirty", False) and self.device.type == "cuda":
            rows = [x for r in self.routes for x in (list(r) + [-1] * (ROUTE_LEN - len(r)))]
            self._dev["routes"] = torch.tensor(rows, dtype=torch.int32, device=self.device)
            self._struct.routes = self._dev["routes"].data_ptr()
            self._struct.NR = len(self.routes)
            self._routes_dirty = False

    def random_spawn_routes(self, R: int, max_spawn: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[R][ML][max_spawn] int32: for every future spawn a random walk as create_random_route draws it (uniform next lane
        at every hop, moving on when the lane was visited; road_network.py:604-646), cut after its first macro lane."""
        out = torch.zeros((R, self.ML, max_spawn), dtype=torch.int32)
        for b in range(R):
            for m, l0 in enumerate(self.micro):
                for k in range(max_spawn):
                    path = [l0]
                    while self.kind[path[-1]] and self.next[path[-1]] and len(path) < ROUTE_LEN:
                        cand = self.next[path[-1]]
                        i0 = int(torch.randint(0, len(cand), (1,), generator=generator))
                        i = i0
                        while cand[i] in path:
                            i = (i + 1) % len(cand)
                            if i == i0:
                                break
                        if cand[i] in path:
                            break
                        path.append(cand[i])
                    out[b, m, k] = self.route_index[tuple(path)]
        return out.to(self.device)

    def default_own(self, R: int, dtype, umax: float) -> torch.Tensor:
        o = torch.zeros((R, self.n_own, 2), dtype=dtype, device=self.device)
        o[..., 1] = umax                      # ARZ.FullQ(u_max), model/macro/_arz.py:59-63
        return o

    def make_aux0(self, R: int, dtype, p0=None, v0=None, a0=None, route0=None, count0=None, capacitor0=None,
                  pid0=None) -> torch.Tensor:
        """Initial ``aux`` rows.  p0, v0, a0 [R][ML][cap]: initial vehicles of every micro lane, HEAD FIRST (the reverse of
        ``lane.curr_vehicle``); route0 [ML][cap] route ids (``route_id``); pid0 [ML][cap] parameter-set ids (rows of
        ``veh_params``; default 0); count0 [ML]; capacitor0 [R][NCAP].  Differentiable wrt p0, v0, a0, capacitor0."""
        dev, ML, cap = self.device, self.ML, self.veh_cap
        z = lambda *s: torch.zeros(s, dtype=dtype, device=dev)
        p0 = z(R, ML, cap) if p0 is None else p0.to(dtype)
        v0 = z(R, ML, cap) if v0 is None else v0.to(dtype)
        a0 = z(R, ML, cap) if a0 is None else a0.to(dtype)
        rid = z(ML, cap) if route0 is None else torch.as_tensor(route0, device=dev).to(dtype)
        pid = z(ML, cap) if pid0 is None else torch.as_tensor(pid0, device=dev).to(dtype)
        cnt = z(ML) if count0 is None else torch.as_tensor(count0, device=dev).to(dtype)
        assert float(cnt.max()) <= cap if ML else True
        capac = z(R, self.NCAP) if capacitor0 is None else capacitor0.to(dtype)
        ex = lambda t: t.reshape(1, -1).expand(R, -1)
        return torch.cat([p0.reshape(R, -1), v0.reshape(R, -1), a0.reshape(R, -1), ex(rid), z(R, ML * cap), ex(pid), z(R, ML), ex(cnt),
                          z(R, ML), capac, z(R, 3)], dim=1).contiguous()


class HybRolloutFn(torch.autograd.Function):
    """(r0, y0, u0, own0, sig, incoming, aux0) -> (hist [T+1,R,4,NC], aux_hist [T+1,R,AUX], head_hist [T,R,ML,2])."""

    @staticmethod
    def forward(ctx, r0, y0, u0, own0, sig, incoming, aux0, topo: HybridNetTopology, route, spawn_route, veh_par, umax, dt,
                steps, soft, flags, ueq0=None, src_rand=None):
        dev = _lib.require_cuda(r0, y0, u0, own0, sig, incoming, aux0, route, spawn_route, flags, src_rand)
        ctx.set_materialize_grads(False)      # an output nobody differentiates stays None in backward (no zero-filled history)
        c = lambda t: None if t is None else t.contiguous()
        r0, y0, u0, own0, sig, incoming, aux0, route, spawn_route = map(c, (r0, y0, u0, own0, sig, incoming, aux0, route, spawn_route))
        R, NC = r0.shape
        assert NC == topo.NC, "state rows must hold the network's macro cells lane by lane"
        dtype, steps, L = r0.dtype, int(steps), topo.L
        assert aux0.shape == (R, topo.AUX) and aux0.dtype == dtype
        if topo.mode == MODE_ITSCP:
            assert sig is not None and incoming is not None and sig.shape == (R, steps, L) == incoming.shape
        per_rep = 0
        if route is not None:
            assert route.dtype == torch.int32 and route.shape[-3:] == (steps, 2, L)
            per_rep = int(route.dim() == 4 and route.shape[0] == R and R > 1)
        assert route is not None or topo.NCAP == 0, "conversions read the step's MacroRoute (road_network.py:129)"
        KS, sp_rep = 0, 0
        if spawn_route is not None:
            assert spawn_route.dtype == torch.int32 and spawn_route.shape[-2] == topo.ML
            KS = int(spawn_route.shape[-1])
            sp_rep = int(spawn_route.dim() == 3 and spawn_route.shape[0] == R and R > 1)
        assert spawn_route is not None or (topo.NCAP == 0 and not topo.sources)
        par = torch.as_tensor(veh_par, dtype=dtype, device=dev).reshape(-1, 6).contiguous()      # [NP, 6] parameter sets
        n_rand, rand_rep = 0, 0
        if topo.sources:
            assert src_rand is not None and src_rand.dtype == dtype, "waiting-list sources consume uniform draws (src_rand)"
            src_rand = src_rand.contiguous()
            n_rand = int(src_rand.shape[-1])
            rand_rep = int(src_rand.dim() == 2 and src_rand.shape[0] == R and R > 1)
        if own0 is None:
            own0 = topo.default_own(R, dtype, umax)
        hist = torch.empty((steps + 1, R, 4, NC), dtype=dtype, device=dev)
        ownh = torch.empty((steps + 1, R, max(topo.n_own, 1), 2), dtype=dtype, device=dev)
        auxh = torch.empty((steps + 1, R, topo.AUX), dtype=dtype, device=dev)
        headh = torch.empty((steps, R, max(topo.ML, 1), 2), dtype=dtype, device=dev)
        lib = _lib.load()
        assert lib.dhts_hyb_aux_size(topo.struct_ptr()) == topo.AUX
        fn = getattr(lib, "dhts_hyb_rollout_fwd_" + suffix(dtype))
        with torch.cuda.device(dev):
            check(fn(topo.struct_ptr(), ptr(topo.real("dx", dtype)), ptr(topo.real("lane_len", dtype)), ptr(route), per_rep,
                     ptr(spawn_route), sp_rep, KS, ptr(sig), ptr(incoming), ptr(par), int(par.shape[0]), creal(dtype, topo.veh_len),
                     ptr(src_rand if topo.sources else None), n_rand, rand_rep, creal(dtype, umax), creal(dtype, dt), steps, R,
                     topo.mode, int(bool(soft)), ptr(r0), ptr(y0), ptr(u0), ptr(c(ueq0)), ptr(own0 if topo.n_own else None), ptr(aux0),
                     ptr(hist), ptr(ownh if topo.n_own else None), ptr(auxh), ptr(headh if topo.ML else None), ptr(flags),
                     stream_ptr(dev)), "dhts_hyb_rollout_fwd")
        ctx.save_for_backward(hist, ownh, auxh, sig, incoming, route, spawn_route, par, src_rand if topo.sources else None)
        ctx.cfg = (topo, per_rep, sp_rep, KS, n_rand, rand_rep, float(umax), float(dt), steps, int(bool(soft)), R)
        ctx.flags = flags
        ctx.mark_non_differentiable(headh)
        return hist, auxh, headh

    @staticmethod
    def backward(ctx, g_hist, g_auxh, _g_head):
        hist, ownh, auxh, sig, incoming, route, spawn_route, par, src_rand = ctx.saved_tensors
        topo, per_rep, sp_rep, KS, n_rand, rand_rep, umax, dt, steps, soft, R = ctx.cfg
        dev, dtype = hist.device, hist.dtype
        NC, L = topo.NC, topo.L
        g_states = g_hist[1:].contiguous() if (g_hist is not None and steps > 0) else None
        g_aux = g_auxh[1:].contiguous() if (g_auxh is not None and steps > 0) else None
        g_r0 = torch.empty((R, NC), dtype=dtype, device=dev); g_y0 = torch.empty_like(g_r0); g_u0 = torch.empty_like(g_r0)
        g_own0 = torch.zeros((R, max(topo.n_own, 1), 2), dtype=dtype, device=dev)
        itscp = topo.mode == MODE_ITSCP
        g_sig = torch.zeros((R, steps, L), dtype=dtype, device=dev) if itscp else None
        g_inc = torch.zeros((R, steps, L), dtype=dtype, device=dev) if itscp else None
        g_aux0 = torch.zeros((R, topo.AUX), dtype=dtype, device=dev)
        fn = getattr(_lib.load(), "dhts_hyb_rollout_bwd_" + suffix(dtype))
        with torch.cuda.device(dev):
            check(fn(topo.struct_ptr(), ptr(topo.real("dx", dtype)), ptr(topo.real("lane_len", dtype)), ptr(route), per_rep,
                     ptr(spawn_route), sp_rep, KS, ptr(sig), ptr(incoming), ptr(par), int(par.shape[0]), creal(dtype, topo.veh_len),
                     ptr(src_rand), n_rand, rand_rep, creal(dtype, umax), creal(dtype, dt), steps, R,
                     topo.mode, soft, ptr(hist), ptr(ownh if topo.n_own else None), ptr(auxh), ptr(g_states), ptr(g_aux),
                     ptr(g_r0), ptr(g_y0), ptr(g_u0), ptr(g_own0 if topo.n_own else None), ptr(g_sig), ptr(g_inc), ptr(g_aux0),
                     ptr(ctx.flags), stream_ptr(dev)), "dhts_hyb_rollout_bwd")
        if g_hist is not None:
            g_r0 = g_r0 + g_hist[0, :, 0]; g_y0 = g_y0 + g_hist[0, :, 1]; g_u0 = g_u0 + g_hist[0, :, 2]
        if g_auxh is not None:      # loss terms on the initial vehicles themselves
            m = torch.zeros(topo.AUX, dtype=dtype, device=dev); m[:topo.A_RID] = 1
            g_aux0 = g_aux0 + g_auxh[0] * m
        need = ctx.needs_input_grad
        out = (g_r0, g_y0, g_u0, g_own0[:, :topo.n_own] if topo.n_own else None, g_sig, g_inc, g_aux0)
        return tuple(g if need[i] else None for i, g in enumerate(out)) + (None,) * 11


class HybridStates:
    """Stored trajectory of a hybrid rollout: views into the kernels' history buffers (all differentiable)."""

    def __init__(self, topo: HybridNetTopology, hist, auxh, headh):
        self.topo, self.hist, self.aux, self.head = topo, hist, auxh, headh

    @property
    def cells(self):
        """[T+1, R, 3, NC] (r, y, u) of every macro cell before each step and after the last."""
        return self.hist[:, :, :3]

    def _veh(self, off):
        t = self.topo
        return self.aux[:, :, off:off + t.ML * t.veh_cap].reshape(self.aux.shape[0], self.aux.shape[1], t.ML, t.veh_cap)

    @property
    def position(self):
        return self._veh(self.topo.A_P)

    @property
    def speed(self):
        return self._veh(self.topo.A_V)

    @property
    def mass(self):
        """Vehicle.a (road/vehicle/vehicle.py:10,18), the ancillary variable that carries the density gradient."""
        return self._veh(self.topo.A_A)

    @property
    def count(self):
        t = self.topo
        return self.aux[:, :, t.A_CNT:t.A_CNT + t.ML].detach().round().long()

    @property
    def front(self):
        t = self.topo
        return self.aux[:, :, t.A_FRONT:t.A_FRONT + t.ML].detach().round().long()

    @property
    def capacitor(self):
        t = self.topo
        return self.aux[:, :, t.A_CAP:t.A_CAP + t.NCAP]

    def order(self):
        """rank [T+1, R, ML, cap]: 0 for the head vehicle, 1 for its follower ...; >= count for free slots."""
        cap = self.topo.veh_cap
        slot = torch.arange(cap, device=self.aux.device).reshape(1, 1, 1, cap)
        return (slot - self.front.unsqueeze(-1)) % cap

    def occupied(self):
        return self.order() < self.count.unsqueeze(-1)

    def by_rank(self):
        """(p, v, a, valid), each [T+1, R, ML, cap] with the vehicles of a lane HEAD FIRST along the last axis."""
        cap = self.topo.veh_cap
        idx = (self.front.unsqueeze(-1) + torch.arange(cap, device=self.aux.device)) % cap
        valid = torch.arange(cap, device=self.aux.device) < self.count.unsqueeze(-1)
        g = lambda x: torch.gather(x, 3, idx)
        return g(self.position), g(self.speed), g(self.mass), valid


def hybrid_rollout(topo: HybridNetTopology, r0, u0, umax: float, dt: float, steps: int, *, sig=None, incoming=None, route=None,
                   spawn_route=None, own0=None, aux0=None, soft: bool = True, veh_params=None,
                   flags: Optional[_lib.Flags] = None, src_rand=None) -> HybridStates:
    """`steps` x RoadNetwork.forward over R replicas of a connected hybrid network.

    r0, u0 [R, NC] density / speed of the macro cells, lane by lane; sig, incoming [R, steps, L] (ITSCP mode);
    route [steps, 2, L] or [R, steps, 2, L] int32 MacroRoute per step (``topo.route_table``); spawn_route [ML, KS] or
    [R, ML, KS] int32 (``topo.random_spawn_routes`` / ``topo.route_id``); aux0 from ``topo.make_aux0`` (default: no
    vehicles, empty capacitors); veh_params [6] or [NP, 6] parameter sets (a_max, a_pref, v_target, s0, T, length), default
    ``default_micro_vehicle(umax)`` -- set 0 is what spawned vehicles get, ``make_aux0(pid0=...)`` names the sets of the
    initial vehicles; src_rand [n] or [R, n]: the uniform draws of the waiting-list sources (``sources=True`` topologies)."""
    flags = flags or _lib.Flags(r0.device)
    R = r0.shape[0]
    if aux0 is None:
        aux0 = topo.make_aux0(R, r0.dtype)
    par = torch.as_tensor(veh_params if veh_params is not None else default_vehicle_params(umax, topo.veh_len),
                          dtype=torch.float64).reshape(-1, 6)
    assert float((par[:, 5] - topo.veh_len).abs().max()) < 1e-12, "one vehicle length per network (road_network.py:60); it is part of the topology"
    ueq = umax * (1.0 - torch.sqrt(torch.clamp(r0, min=0.0) + EPSILON))
    y0 = r0 * (u0 - ueq)                           # set_r_u, _arz.py:82-86 (autograd: true derivative)
    hist, auxh, headh = HybRolloutFn.apply(r0, y0, u0, own0, sig, incoming, aux0, topo, route, spawn_route, par, float(umax),
                                           float(dt), int(steps), bool(soft), flags.t, None, src_rand)
    return HybridStates(topo, hist, auxh, headh)
