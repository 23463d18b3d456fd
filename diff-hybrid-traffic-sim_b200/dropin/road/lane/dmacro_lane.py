"""dMacroLane / dMacroForwardLayer of the drop-in API (reference: road/lane/dmacro_lane.py).

The reference's layer detaches the lane to Python floats, steps it, and stores an fp32
Jacobian band per step for backward (:96-132, :234-310).  Here the layer is a thin shim
over ``dhts_b200.ops.ArzStepFn``: forward = dhts_arz_step_fwd, backward = dhts_arz_step_bwd
recomputed from the saved inputs; no band is kept (``d_lane`` stays empty).
"""
from typing import List

from road.lane._macro_lane import MacroLane


class dMacroLane(MacroLane):
    class dLane:
        """Kept for interface compatibility: this build never materialises the Jacobian band."""

        def __init__(self, num_cell):
            self.dqs = None

    def __init__(self, id: int, lane_length: float, speed_limit: float, cell_length: float):
        super().__init__(id, lane_length, speed_limit, cell_length)
        self.d_lane: List[dMacroLane.dLane] = []
        self.b_curr_cell: List[MacroLane.Cell] = []

    def vectorize_input(self):
        """(r, y) of (left ghost, cells, right ghost): the differentiable inputs of the layer."""
        return self._padded("r"), self._padded("y")

    def forward(self, delta_time: float):
        cr, cy = self.vectorize_input()
        nr, ny = dMacroForwardLayer.apply(self, cr, cy, delta_time)
        from model.macro._arz import ARZ
        self._assign("next", nr, ny, self._last_nu, ARZ.compute_u_eq(nr, self.speed_limit))

    def clear(self):
        super().clear()
        self.b_curr_cell.clear()


class dMacroForwardLayer:
    """Same call contract as the reference's autograd.Function: ``apply(lane, r[N+2], y[N+2], dt) ->
    (nr[N], ny[N])``, gradients flow to every entry of r and y including the two ghost entries."""

    @staticmethod
    def apply(lane: dMacroLane, r, y, delta_time: float):
        nr, ny, lane._last_nu = lane._step(r, y, delta_time)
        return nr, ny
