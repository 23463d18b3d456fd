"""Macro <-> micro boundary exchange of the drop-in API (reference: road/network/conversion.py:15-215).

The event tests stay on the host (they decide list surgery on ``curr_vehicle``); the
differentiable arithmetic -- flux capacitor / spawn state, density deposition of an absorbed
vehicle -- runs in dhts_m2c_* / dhts_c2m_* (csrc/convert_kernels.cu) with J = 1 junction.
"""
import torch

from dhts_b200 import functional as F
from dhts_b200.dropin import runtime as rt
from road.lane._macro_lane import MacroLane
from road.lane._micro_lane import MicroLane, MicroVehicle
from road.network.route import MicroRoute  # noqa: F401  (re-exported like the reference module)


def _one(x, dtype):
    return rt.scalar(x, dtype).reshape(1)


class Conversion:
    @staticmethod
    def macro_to_macro(network, prev_lane: MacroLane, next_lane: MacroLane):
        """Macro lanes exchange through their ghost cells (RoadNetwork.get_macro_boundary); nothing to do here."""

    @staticmethod
    def macro_to_micro(network, prev_lane: MacroLane, next_lane: MicroLane, delta_time: float):
        """Charge the flux capacitor of the junction with r_last u_last dt; once it holds one vehicle length
        and the micro lane has that much free room at its entrance, emit a default vehicle at position 0 with
        the last cell's speed.  Its `a` equals the vehicle length and carries the capacitor's gradient; the
        capacitor restarts from the remainder with its history cut (conversion.py:32-68)."""
        sd, st = rt.step_dtype(), rt.store_dtype()
        last = prev_lane.curr_cell[-1].state
        nv = MicroVehicle.default_micro_vehicle(next_lane.speed_limit)
        free = next_lane.entering_free_space()
        free = free.detach() if rt.is_tensor(free) else free
        cap_out, spawn, v_new, a_new = F.macro_to_micro(
            _one(prev_lane.flux_capacitor.get(next_lane.id, 0.0), sd), _one(last.q.r, sd), _one(last.u, sd),
            _one(free, sd), _one(nv.length, sd), float(delta_time))
        prev_lane.flux_capacitor[next_lane.id] = cap_out[0].to(st)
        if int(spawn.item()):
            nv.position = 0
            nv.speed = v_new[0].to(st)
            nv.a = a_new[0].to(st)
            network.add_vehicle(nv, network.create_random_route(next_lane.id))

    @staticmethod
    def micro_to_macro(network, prev_lane: MicroLane):
        """Once the head vehicle is a full length past the lane end it leaves the micro lane and its `a`
        is deposited as density into the downstream cells it overlaps; those cells take its speed
        (conversion.py:99-171)."""
        if not prev_lane.num_vehicle():
            return
        hv = prev_lane.get_head_vehicle()
        next_lane: MacroLane = network.lane[network.micro_route[hv.id].next_lane_id()]
        assert next_lane.is_macro(), ""
        if not (hv.position > prev_lane.length + 1.0 * hv.length):
            return
        prev_lane.curr_vehicle = prev_lane.curr_vehicle[:-1]
        sd, st = rt.step_dtype(), rt.store_dtype()
        r, y, u, e = next_lane.device_state()
        row = lambda t: t.to(sd).unsqueeze(0)
        r2, y2, u2, absorbed, _ = F.micro_to_macro(
            _one(hv.position, sd), _one(hv.speed, sd), _one(hv.a, sd), _one(hv.length, sd), _one(prev_lane.length, sd),
            row(r), row(y), row(u), _one(next_lane.cell_length, sd), _one(next_lane.speed_limit, sd))
        next_lane._assign("curr", r2[0].to(st), y2[0].to(st), u2[0].to(st), e)     # stored u_eq stays as it was

    @staticmethod
    def micro_to_micro(network, prev_lane: MicroLane):
        """Head vehicle past the lane end moves to the tail of the next lane on its route."""
        if not prev_lane.num_vehicle():
            return
        hv = prev_lane.get_head_vehicle()
        hr = network.micro_route[hv.id]
        next_lane: MicroLane = network.lane[hr.next_lane_id()]
        assert next_lane.is_micro(), ""
        if hv.position >= prev_lane.length:
            prev_lane.curr_vehicle = prev_lane.curr_vehicle[:-1]
            hv.position = hv.position - prev_lane.length
            next_lane.add_tail_vehicle(hv)
            hr.increment_curr_idx()

    @staticmethod
    def micro_to_none(network, prev_lane: MicroLane):
        """Head vehicle past the end of a lane with no successor leaves the network."""
        if prev_lane.num_vehicle() and prev_lane.get_head_vehicle().position >= prev_lane.length:
            prev_lane.curr_vehicle = prev_lane.curr_vehicle[:-1]
