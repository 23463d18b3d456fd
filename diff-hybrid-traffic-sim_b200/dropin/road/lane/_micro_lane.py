"""Microscopic (IDM) lane of the drop-in API; the step runs in dhts_idm_step_{fwd,bwd}_*.

Object surface of the reference's road/lane/_micro_lane.py: an ordered, caller-mutable list
of ``MicroVehicle`` records (index 0 = tail, the leader of i is i+1), a ghost leader for the
head given by (head_position_delta, head_speed_delta), Euler step into ``next_vehicle_*``.
``forward`` is one launch of the batched CSR kernel with a single lane; positions and speeds
handed to the vehicles are 0-dim views of the device vectors (see dropin/runtime.py).
"""
from typing import List

import torch

from dhts_b200 import functional as F
from dhts_b200.dropin import runtime as rt
from dmath.operation import sigmoid
from road.lane._base_lane import BaseLane, synced
from road.vehicle.micro_vehicle import MicroVehicle

DEFAULT_HEAD_POSITION_DELTA = 1000
DEFAULT_HEAD_SPEED_DELTA = 0
POSITION_DELTA_EPS = 1e-5


class MicroLane(BaseLane):
    # state a queued network step would change: reads and writes run the queue first (dropin/deferred.py)
    curr_vehicle = synced("_curr_vehicle")
    next_vehicle_position = synced("_next_vehicle_position")
    next_vehicle_speed = synced("_next_vehicle_speed")

    def __init__(self, id: int, lane_length: float, speed_limit: float):
        super().__init__(id, lane_length, speed_limit)
        self.curr_vehicle: List[MicroVehicle] = []
        self.acc_info = []                    # per-vehicle clip / collision bits of the last step when record_flags
        self.record_flags = False
        self.next_vehicle_position = []
        self.next_vehicle_speed = []
        self.head_position_delta = DEFAULT_HEAD_POSITION_DELTA
        self.head_speed_delta = DEFAULT_HEAD_SPEED_DELTA
        self._curr_views = None               # (Views position, Views speed) last handed to curr_vehicle
        self._next_views = None
        self._param_cache = None              # (vehicles, params[6,n], lane_off, veh_lane)
        self._head_cache = None               # ((dp, dv), tensor[1,2]) when both are plain numbers

    def is_macro(self):
        return False

    def is_micro(self):
        return True

    # ------------------------------------------------------------------ vehicle list
    def num_vehicle(self):
        return len(self.curr_vehicle)

    def add_head_vehicle(self, vehicle: MicroVehicle):
        self.curr_vehicle.append(vehicle)

    def add_tail_vehicle(self, vehicle: MicroVehicle):
        self.curr_vehicle.insert(0, vehicle)

    def add_vehicle(self, vehicle: MicroVehicle):
        """Place a vehicle by position.  As in the reference (_micro_lane.py:61-113) a vehicle may only enter
        behind the tail or ahead of the head, and must clear half the summed lengths of its neighbour."""
        assert vehicle.position >= 0 and vehicle.position <= self.length, ""
        gap = lambda a, b: (a.length + b.length) * 0.5
        if not self.curr_vehicle:
            self.curr_vehicle.append(vehicle)
        elif self.curr_vehicle[0].position > vehicle.position:
            assert self.curr_vehicle[0].position - vehicle.position >= gap(self.curr_vehicle[0], vehicle), ""
            self.curr_vehicle.insert(0, vehicle)
        else:
            for v in self.curr_vehicle:
                assert not (v.position > vehicle.position), ""      # would be an insertion between two vehicles
                assert vehicle.position - v.position > gap(vehicle, v), ""
            self.curr_vehicle.append(vehicle)

    def get_head_vehicle(self):
        assert self.curr_vehicle, ""
        return self.curr_vehicle[-1]

    def get_tail_vehicle(self):
        assert self.curr_vehicle, ""
        return self.curr_vehicle[0]

    # ------------------------------------------------------------------ vehicles <-> device vectors
    def _state(self):
        cv = self._curr_views
        if cv is not None and cv[0].still(v.position for v in self.curr_vehicle) \
                and cv[1].still(v.speed for v in self.curr_vehicle):
            return cv[0].vec, cv[1].vec
        return rt.gather([v.position for v in self.curr_vehicle]), rt.gather([v.speed for v in self.curr_vehicle])

    def _hand_out(self, p, v):
        vp, vv = rt.Views(p), rt.Views(v)
        for mv, a, b in zip(self.curr_vehicle, vp.items, vv.items):
            mv.position, mv.speed = a, b
        self._curr_views = (vp, vv)

    def _params(self, dtype):
        """[6, n] IDM parameters of the vehicles on the lane.  The reference re-reads the six attributes of every
        vehicle at every step (_micro_lane.py:195-214), so the cache is keyed on their VALUES (a tuple per vehicle),
        not only on which vehicle objects are on the lane: an in-place edit of e.g. target_speed takes effect."""
        c = self._param_cache
        veh = self.curr_vehicle
        vals = [tuple(mv.idm_params()) for mv in veh]
        if c is None or c[4] != vals or any(a is not b for a, b in zip(c[0], veh)) or c[1].dtype != dtype:
            n = len(veh)
            par = torch.tensor(vals, dtype=dtype).reshape(n, 6).t().contiguous().to(rt.device())
            off = torch.tensor([0, n], dtype=torch.int32, device=rt.device())
            c = self._param_cache = (list(veh), par, off, torch.zeros(max(n, 1), dtype=torch.int32, device=rt.device()), vals)
        return c[1], c[2], c[3]

    def _head(self, dtype):
        dp, dv = self.head_position_delta, self.head_speed_delta
        if not rt.is_tensor(dp) and not rt.is_tensor(dv):
            c = self._head_cache
            if c is None or c[0] != (dp, dv) or c[1].dtype != dtype:
                c = self._head_cache = ((dp, dv), torch.tensor([[float(dp), float(dv)]], dtype=dtype).to(rt.device()))
            return c[1]
        return rt.gather([dp, dv], dtype).reshape(1, 2)

    # ------------------------------------------------------------------ the step
    def _step(self, p, v, head, delta_time):
        """(p, v)[n], head[1,2] -> (np, nv)[n]; collisions are flagged on the device, deltas zeroed, and
        reported print-and-continue by runtime.check_flags (reference: _micro_lane.py:149-166)."""
        sd, st = rt.step_dtype(), rt.store_dtype()
        par, off, veh_lane = self._params(sd)
        out = F.idm_step(p.to(sd), v.to(sd), par, off, head.to(sd), float(delta_time), rt.flags(), veh_lane=veh_lane,
                         want_flags=self.record_flags)
        if self.record_flags:
            bits = out[2].tolist()
            self.acc_info = [(None, None, bool(b & 1), bool(b & 2)) for b in bits]
        return out[0].to(st), out[1].to(st)

    def forward(self, delta_time: float):
        """Next positions / speeds into `next_vehicle_*`; `update_state` applies them."""
        if not self.curr_vehicle:
            self.next_vehicle_position, self.next_vehicle_speed, self._next_views = [], [], None
            return
        p, v = self._state()
        np_, nv_ = self._step(p, v, self._head(rt.step_dtype()), delta_time)
        self._set_next(np_, nv_)

    def _set_next(self, np_, nv_):
        self._next_views = (rt.Views(np_), rt.Views(nv_))
        self.next_vehicle_position = list(self._next_views[0].items)
        self.next_vehicle_speed = list(self._next_views[1].items)

    def compute_state_delta(self, id):
        """Gap and closing speed to the leader as the step sees them (host copy, for callers that inspect it)."""
        if id == len(self.curr_vehicle) - 1:
            return self.head_position_delta, self.head_speed_delta
        mv, lv = self.curr_vehicle[id], self.curr_vehicle[id + 1]
        return abs(lv.position - mv.position) - (lv.length + mv.length) * 0.5, mv.speed - lv.speed

    def update_state(self):
        for i, mv in enumerate(self.curr_vehicle):
            mv.position, mv.speed = self.next_vehicle_position[i], self.next_vehicle_speed[i]
        nv = self._next_views
        ok = nv is not None and nv[0].still(self.next_vehicle_position) and nv[1].still(self.next_vehicle_speed)
        self._curr_views = nv if ok else None

    # ------------------------------------------------------------------ vector access (host tensors out)
    def set_state_vector(self, position, speed):
        assert len(position) == self.num_vehicle() and len(speed) == self.num_vehicle(), "Vehicle number mismatch"
        self._hand_out(rt.vector(position), rt.vector(speed))

    def get_state_vector(self):
        rt.check_flags()
        if not self.curr_vehicle:
            z = torch.zeros((0,), dtype=rt.store_dtype())
            return z, z.clone()
        p, v = self._state()
        return p.cpu(), v.cpu()

    def set_next_state_vector(self, position, speed):
        assert len(position) == self.num_vehicle() and len(speed) == self.num_vehicle(), "Vehicle number mismatch"
        self._set_next(rt.vector(position), rt.vector(speed))

    def get_next_state_vector(self):
        rt.check_flags()
        if not self.next_vehicle_position:
            z = torch.zeros((0,), dtype=rt.store_dtype())
            return z, z.clone()
        return rt.gather(self.next_vehicle_position).cpu(), rt.gather(self.next_vehicle_speed).cpu()

    def device_state(self):
        """(position, speed) device vectors, no host copy."""
        return self._state()

    # ------------------------------------------------------------------ helpers of the hybrid exchange
    def entering_free_space(self):
        if self.curr_vehicle:
            t = self.curr_vehicle[0]
            return t.position - 0.5 * t.length
        return self.length

    def on_this_lane(self, position, differentiable: bool):
        if not rt.is_tensor(position):
            position = torch.tensor(position)
        if differentiable:
            return sigmoid(position, constant=16.0) * sigmoid(self.length - position, constant=16.0)
        return float(position >= 0 and position <= self.length)

    def clear(self):
        self.curr_vehicle.clear()
        self.next_vehicle_position, self.next_vehicle_speed = [], []
        self._curr_views = self._next_views = None
