"""dMicroLane / dMicroForwardLayer of the drop-in API (reference: road/lane/dmicro_lane.py).

The reference stores a per-step fp32 band of 2x2 Jacobians (dIDM, :87-127) and applies it in
backward (:271-297).  Here forward = dhts_idm_step_fwd and backward = dhts_idm_step_bwd
recomputed from the saved inputs through ``dhts_b200.ops.IdmStepFn``; the ghost leader of
``vectorize_input`` is folded into the kernel and its adjoint comes back as the gradient of
(head_position_delta, head_speed_delta).
"""
from typing import List

import torch

from dhts_b200.dropin import runtime as rt
from road.lane._micro_lane import MicroLane
from road.vehicle.micro_vehicle import MicroVehicle


class dMicroLane(MicroLane):
    class dLane:
        """Kept for interface compatibility: this build never materialises the Jacobian band."""

        def __init__(self, num_vehicle):
            self.dqs = None

    def __init__(self, id: int, lane_length: float, speed_limit: float):
        super().__init__(id, lane_length, speed_limit)
        self.d_lane: List[dMicroLane.dLane] = []
        self.b_curr_vehicle: List[MicroVehicle] = []

    def vectorize_input(self):
        """(positions, speeds) of the vehicles followed by the ghost leader (p_head + dp, v_head - dv)."""
        if not self.curr_vehicle:
            z = torch.zeros((0,), dtype=rt.store_dtype(), device=rt.device())
            return z, z.clone()
        p, v = self._state()
        h = self._head(p.dtype)[0]
        return torch.cat([p, (p[-1] + h[0]).reshape(1)]), torch.cat([v, (v[-1] - h[1]).reshape(1)])

    def clear_gradient(self):
        self.d_lane = []

    def clear(self):
        super().clear()
        self.b_curr_vehicle.clear()


class dMicroForwardLayer:
    """Same call contract as the reference's autograd.Function: ``apply(lane, p[n+1], s[n+1], dt) ->
    (np[n], ns[n])`` with the ghost leader as last entry; gradients reach all n+1 entries."""

    @staticmethod
    def apply(lane: dMicroLane, p, s, delta_time: float):
        sd = rt.step_dtype()
        p, s = p.to(sd), s.to(sd)
        head = torch.stack([p[-1] - p[-2], s[-2] - s[-1]]).reshape(1, 2)
        return lane._step(p[:-1], s[:-1], head, delta_time)
