"""Network step orchestration of the drop-in API (reference: road/network/road_network.py).

``forward`` keeps the reference's order -- boundaries of every lane, ``forward`` of every
lane, ``update_state`` of every lane (Jacobi), then the macro<->micro conversions in lane-id
order (:79-127) -- with each lane step and each conversion running on the GPU.  The
boundary rules (ghost-cell source of a macro lane, cross-lane leader of a micro lane's head
vehicle) are host logic over a handful of scalars and are restated here; subclasses such as
ItscpRoadNetwork override the same three hooks as in the reference.
"""
from typing import Dict

import numpy as np

from dhts_b200.dropin import deferred
from dhts_b200.dropin import runtime as rt
from road.lane._base_lane import BaseLane
from road.lane._macro_lane import MacroLane
from road.lane._micro_lane import DEFAULT_HEAD_POSITION_DELTA, DEFAULT_HEAD_SPEED_DELTA, MicroLane
from road.network.conversion import Conversion
from road.network.route import MacroRoute, MicroRoute
from road.vehicle.micro_vehicle import DEFAULT_VEHICLE_LENGTH, MicroVehicle

MAX_ROUTE_LENGTH = 32


def _synced(name):
    """Network attribute a queued step would change: reads and writes run the queue first (dropin/deferred.py)."""
    def get(self):
        if self.__dict__.get("_pending"):
            self.flush()
        return self.__dict__[name]

    def put(self, value):
        if self.__dict__.get("_pending"):
            self.flush()
        self.__dict__[name] = value
    return property(get, put)


class RoadNetwork:
    vehicle = _synced("_vehicle")
    micro_route = _synced("_micro_route")
    macro_route = _synced("_macro_route")
    num_vehicle = _synced("_num_vehicle")

    def __init__(self, speed_limit: float):
        self._pending = 0                 # forward() calls queued for one fused rollout launch (dropin/deferred.py)
        self._pending_dt = self._pending_diff = self._pending_mode = None
        self._defer_cache = {}
        self._stepped_once = False
        self.lane: Dict[int, BaseLane] = {}
        self.speed_limit = speed_limit                  # one limit for the whole network
        self.vehicle_length = DEFAULT_VEHICLE_LENGTH    # one vehicle length for the whole network
        self.macro_route = MacroRoute()
        self.vehicle: Dict[int, MicroVehicle] = {}
        self.micro_route: Dict[int, MicroRoute] = {}
        self.num_lane = 0
        self.num_vehicle = 0

    # ------------------------------------------------------------------ construction
    def add_lane(self, lane: BaseLane):
        assert isinstance(lane, (MacroLane, MicroLane)), ""
        self.flush()
        lane.speed_limit = self.speed_limit
        lane.id = self.num_lane
        lane._net = self
        self.lane[lane.id] = lane
        self.num_lane += 1
        return lane.id

    def add_vehicle(self, nv: MicroVehicle, route: MicroRoute):
        assert nv.length == self.vehicle_length, ""
        nv.id = self.num_vehicle
        self.num_vehicle += 1
        self.vehicle[nv.id] = nv
        self.micro_route[nv.id] = route
        lane = self.lane[route.curr_lane_id()]
        assert lane.is_micro()
        lane.add_vehicle(nv)
        return nv.id

    def connect_lane(self, prev_lane_id: int, next_lane_id: int):
        self.flush()
        a, b = self.lane[prev_lane_id], self.lane[next_lane_id]
        a.add_next_lane(b)
        b.add_prev_lane(a)

    # ------------------------------------------------------------------ the step
    def forward(self, delta_time: float, differentiable: bool):
        """One step of the whole network.  The step is QUEUED when the network is one the fused rollout kernels can take
        (dropin/deferred.py) and runs, together with the steps queued after it, when its result is first looked at."""
        mode = deferred.plan(self)
        if mode is None or not self._stepped_once:
            # also the very first step of a network runs at once: a time step that violates the CFL condition then raises
            # here, inside forward(), as in the reference (_macro_lane.py:141-146); later ones raise at the flush
            self.flush()
            self._stepped_once = True
            return self._forward_now(delta_time, differentiable)
        if self._pending and (delta_time != self._pending_dt or differentiable != self._pending_diff
                              or mode != self._pending_mode):
            self.flush()
        self._pending_dt, self._pending_diff, self._pending_mode = delta_time, differentiable, mode
        self._pending += 1

    def flush(self):
        """Run the queued steps now."""
        if self._pending:
            deferred.flush(self)

    def _forward_now(self, delta_time: float, differentiable: bool):
        lanes = list(self.lane.values())
        for lane in lanes:
            self.setup_boundary(lane.id, differentiable)
        for lane in lanes:
            lane.forward(delta_time)
        for lane in lanes:
            lane.update_state()
        self.conversion(delta_time)
        rt.check_flags()        # CFL assert / collision report of this step (reference raises inside the step)

    def conversion(self, delta_time: float):
        for lane in list(self.lane.values()):
            if lane.is_macro():
                self.conversion_macro(lane, delta_time)
            else:
                self.conversion_micro(lane)

    def conversion_macro(self, lane: MacroLane, delta_time: float):
        nid = self.macro_route.get_next_lane(lane.id)
        if nid == -1:
            return
        nxt = self.lane[nid]
        if nxt.is_macro():
            Conversion.macro_to_macro(self, lane, nxt)
        else:
            Conversion.macro_to_micro(self, lane, nxt, delta_time)

    def conversion_micro(self, lane: MicroLane):
        if not lane.num_vehicle():
            return
        nid = self.micro_route[lane.get_head_vehicle().id].next_lane_id()
        if nid == -1:
            Conversion.micro_to_none(self, lane)
        elif self.lane[nid].is_macro():
            Conversion.micro_to_macro(self, lane)
        else:
            Conversion.micro_to_micro(self, lane)

    def setup_boundary(self, id: int, differentiable: bool):
        if self.lane[id].is_macro():
            self.setup_macro_boundary(id, differentiable)
        else:
            self.setup_micro_boundary(id, differentiable)

    # ------------------------------------------------------------------ macro boundaries
    def get_macro_boundary(self, id: int, left: bool, differentiable: bool):
        """(r, u) of the ghost cell on one side: the facing edge cell of the adjacent MACRO lane (the only
        neighbour, or the one the current MacroRoute selects among several), else the lane's own ghost cell."""
        lane: MacroLane = self.lane[id]
        adj = lane.prev_lane if left else lane.next_lane
        cell = lane.get_leftmost_cell() if left else lane.get_rightmost_cell()
        if len(adj) == 1:
            other = next(iter(adj.values()))
        elif len(adj) > 1:
            other = self.lane[self.macro_route.get_prev_lane(id) if left else self.macro_route.get_next_lane(id)]
        else:
            other = None
        if other is not None and other.is_macro():
            cell = other.curr_cell[-1] if left else other.curr_cell[0]
        return cell.state.q.r, cell.state.u

    def setup_macro_boundary(self, id: int, differentiable: bool):
        lane: MacroLane = self.lane[id]
        assert lane.is_macro(), ""
        lane.set_leftmost_cell(*self.get_macro_boundary(id, True, differentiable))
        lane.set_rightmost_cell(*self.get_macro_boundary(id, False, differentiable))

    def get_macro_state_of_micro_lane(self, id: int, differentiable: bool):
        """(density, mean speed) a micro lane would show as a macro cell, counting vehicles about to enter from
        micro predecessors and vehicles that just left into micro successors (soft membership when differentiable)."""
        lane: MicroLane = self.lane[id]
        assert lane.is_micro(), ""
        seen = [(v, v.position) for v in lane.curr_vehicle]
        for p in lane.prev_lane.values():
            if p.is_micro():
                seen += [(v, -(p.length - v.position)) for v in p.curr_vehicle
                         if self.micro_route[v.id].next_lane_id() == id]
        for n in lane.next_lane.values():
            if n.is_micro():
                seen += [(v, lane.length + v.position) for v in n.curr_vehicle
                         if self.micro_route[v.id].prev_lane_id() == id]
        density, speed_sum, count = 0, 0, 0
        for v, pos in seen:
            w = lane.on_this_lane(pos, differentiable)
            density = density + w * (v.length / lane.length)
            speed_sum = speed_sum + w * v.speed
            count = count + w
        density = min(density, 1.0)
        return density, (speed_sum / count if count > 0 else self.speed_limit)

    def create_random_macro_route(self):
        """One successor per macro lane, each lane fed by at most one predecessor, drawn with np.random in the
        reference's call order (two permutations per macro lane) so that seeded runs agree."""
        route = MacroRoute()
        for lane_id in np.random.permutation(list(self.lane.keys())):
            lane = self.lane[lane_id]
            if lane.is_micro():
                continue
            for nid in np.random.permutation(list(lane.next_lane.keys())):
                if nid not in route.prev_lane_dict:
                    route.next_lane_dict[lane_id] = nid
                    route.prev_lane_dict[nid] = lane_id
                    break
        return route

    # ------------------------------------------------------------------ micro boundaries
    def setup_micro_boundary(self, id: int, differentiable: bool):
        """Ghost leader of the head vehicle: walk its route; the first micro lane on it that holds vehicles
        supplies its tail vehicle as leader (gap = rest of this lane + empty lanes in between + the leader's
        offset, floored at 0); a macro lane on the route, or no leader at all, gives the defaults."""
        lane: MicroLane = self.lane[id]
        assert lane.is_micro(), ""
        lane.head_position_delta = DEFAULT_HEAD_POSITION_DELTA
        lane.head_speed_delta = DEFAULT_HEAD_SPEED_DELTA
        if lane.num_vehicle() == 0:
            return
        hv: MicroVehicle = self.vehicle[lane.get_head_vehicle().id]
        hr: MicroRoute = self.micro_route[hv.id]
        ahead = lane.length - hv.position - hv.length * 0.5
        for k in range(hr.curr_idx, hr.route_length() - 1):
            here, nxt = self.lane[hr.route[k]], self.lane[hr.route[k + 1]]
            link = here.next_lane.get(nxt.id)
            if link is not None and link is not nxt:
                link = None
            if isinstance(link, MacroLane):
                return
            if isinstance(link, MicroLane) and link.num_vehicle():
                lv = link.get_tail_vehicle()
                gap = ahead + (lv.position - lv.length * 0.5)
                gap = 0.0 if 0.0 > gap else gap
                # single on-route leader, weight 1 (the reference's score-weighted mean degenerates to this)
                lane.head_position_delta = 0 + 1.0 * gap
                lane.head_speed_delta = 0 + 1.0 * (hv.speed - lv.speed)
                return
            if link is not None and not isinstance(link, (MacroLane, MicroLane)):
                raise ValueError()
            ahead = ahead + nxt.length

    def create_default_vehicle_with_random_route(self, lane_id: int):
        return MicroVehicle.default_micro_vehicle(self.speed_limit), self.create_random_route(lane_id)

    def create_random_vehicle_with_random_route(self, lane_id: int):
        return MicroVehicle.random_micro_vehicle(self.speed_limit), self.create_random_route(lane_id)

    def create_random_route(self, lane_id: int):
        """Random walk over successors (one np.random.randint per hop, as the reference), preferring lanes not
        yet on the route, at most MAX_ROUTE_LENGTH lanes."""
        route, cur = [], lane_id
        for _ in range(MAX_ROUTE_LENGTH):
            route.append(cur)
            lane = self.lane[cur]
            if not lane.has_next_lane():
                break
            succ = list(lane.next_lane.keys())
            first = int(np.random.randint(0, len(succ)))
            pick = first
            while succ[pick] in route:
                pick = (pick + 1) % len(succ)
                if pick == first:
                    break
            cur = succ[pick]
        return MicroRoute(route)
