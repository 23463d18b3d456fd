"""Macroscopic (ARZ) lane of the drop-in API; the step runs in dhts_arz_step_{fwd,bwd}_*.

Object surface of the reference's road/lane/_macro_lane.py (cells with assignable
``state.q.r / q.y / u / u_eq``, ghost cells, flux capacitor, vector get/set), but the cells
are views of four device vectors and ``forward`` is ONE kernel launch:

    reference                                      here
    _solve_riemann + update loop  (:83-146)        dhts_arz_step_fwd  (csrc/arz_kernels.cu)
    set_next_state_vector_y's u   (:282-299)       `nu` output of the same launch
    dARZ Jacobian band + VJP (dmacro_lane.py)      dhts_arz_step_bwd via autograd

Both ``MacroLane`` and ``dMacroLane`` differentiate with the reference's ANALYTIC adjoint
(SURVEY.md App. B.1b); the reference's pure-autodiff variant is not reproduced.
"""
import math
from typing import Dict, List

import torch

from dhts_b200 import functional as F
from dhts_b200.dropin import runtime as rt
from model.macro._arz import ARZ
from road.lane._base_lane import BaseLane, synced

_FIELDS = ("r", "y", "u", "e")


def _get(cell, k):
    s = cell.state
    return s.q.r if k == "r" else s.q.y if k == "y" else s.u if k == "u" else s.u_eq


class MacroLane(BaseLane):
    class Cell:
        """[start, end) stretch of the lane with its ARZ record."""

        def __init__(self, start, end, speed_limit):
            self.start, self.end = start, end
            self.state = ARZ.FullQ(speed_limit)

    # state a queued network step would change: reads and writes run the queue first (dropin/deferred.py)
    curr_cell = synced("_curr_cell")
    next_cell = synced("_next_cell")
    flux_capacitor = synced("_flux_capacitor")

    def __init__(self, id: int, lane_length: float, speed_limit: float, cell_length: float):
        super().__init__(id, lane_length, speed_limit)
        self.num_cell = math.ceil(self.length / cell_length)
        assert self.num_cell > 0, "Number of cells in a road must be larger than 0."
        dx = self.cell_length = self.length / self.num_cell
        mk = lambda: [MacroLane.Cell(dx * i, dx * (i + 1), speed_limit) for i in range(self.num_cell)]
        self.curr_cell: List[MacroLane.Cell] = mk()
        self.next_cell: List[MacroLane.Cell] = mk()
        # ghost cells used when no macro lane is connected on that side; no flow by default
        self.leftmost_cell = MacroLane.Cell(0, 0, speed_limit)
        self.rightmost_cell = MacroLane.Cell(dx, dx, speed_limit)
        self.riemann_solution = None          # the solver lives in the kernel; see `riemann_case`
        self.riemann_case = None              # int32 [N+1] outcome per interface when record_case is set
        self.record_case = False
        self.flux_capacitor: Dict[int, float] = {}
        self.bdry_callback = None
        self.bdry_callback_args = {"lane": self}
        self._views = {"curr": None, "next": None}      # field -> rt.Views of the device vectors
        self._ghost_src = {}                            # side -> (r, u) objects the ghost was derived from
        self._ghost_dev = {}                            # (cell id, field) -> (source object, [1] device tensor, (dtype, device))

    def is_macro(self):
        return True

    def is_micro(self):
        return False

    # ------------------------------------------------------------------ cells <-> device vectors
    def _assign(self, which, r, y, u, e):
        cells = self.curr_cell if which == "curr" else self.next_cell
        v = {"r": rt.Views(r), "y": rt.Views(y), "u": rt.Views(u), "e": rt.Views(e)}
        R, Y, U, E = v["r"].items, v["y"].items, v["u"].items, v["e"].items
        for i, c in enumerate(cells):
            s = c.state
            s.q = ARZ.Q(R[i], Y[i])
            s.u, s.u_eq, s.u_max = U[i], E[i], self.speed_limit
        self._views[which] = v

    def _vec(self, which, k):
        """Device vector of field k; the cached one unless somebody rewrote a cell attribute."""
        cells = self.curr_cell if which == "curr" else self.next_cell
        v = self._views[which]
        if v is not None and v[k].still(_get(c, k) for c in cells):
            return v[k].vec
        return rt.gather([_get(c, k) for c in cells])

    def _ghost(self, cell, k):
        """[1] device tensor of a ghost-cell field; cached while the cell keeps the very same constant object (static ghosts
        are set once, road_network.py:312-321), so a step does not pay eight host->device copies for them."""
        v = _get(cell, k)
        if rt.is_tensor(v) and v.requires_grad:
            return rt.scalar(v).reshape(1)
        key = (id(cell), k)
        hit = self._ghost_dev.get(key)
        if hit is not None and hit[0] is v and hit[2] == (rt.store_dtype(), rt.device()):
            return hit[1]
        t = rt.scalar(v).reshape(1)
        self._ghost_dev[key] = (v, t, (rt.store_dtype(), rt.device()))
        return t

    def _padded(self, k, detach=False):
        """[N+2] = (left ghost, cells, right ghost), the operator's input layout (dmacro_lane.py:134-158)."""
        p = torch.cat([self._ghost(self.leftmost_cell, k), self._vec("curr", k), self._ghost(self.rightmost_cell, k)])
        return p.detach() if detach else p

    # ------------------------------------------------------------------ the step
    def _step(self, cr, cy, delta_time):
        """(r, y)[N+2] -> (nr, ny, nu)[N]: one launch of the batched kernel with B = 1."""
        sd, st = rt.step_dtype(), rt.store_dtype()
        row = lambda t: t.to(sd).unsqueeze(0)
        out = F.arz_step(row(cr), row(cy), row(self._padded("u", True)), float(self.cell_length),
                         float(self.speed_limit), float(delta_time), rt.flags(), ueq_pad=row(self._padded("e", True)),
                         want_case=self.record_case)
        if self.record_case:
            self.riemann_case = out[3][0]
        return out[0][0].to(st), out[1][0].to(st), out[2][0].to(st)

    def forward(self, delta_time: float):
        """Next state into `next_cell`; `update_state` applies it."""
        nr, ny, nu = self._step(self._padded("r"), self._padded("y"), delta_time)
        self._assign("next", nr, ny, nu, ARZ.compute_u_eq(nr, self.speed_limit))

    def update_state(self):
        for c, n in zip(self.curr_cell, self.next_cell):
            c.state.q.r, c.state.q.y, c.state.u, c.state.u_eq = n.state.q.r, n.state.q.y, n.state.u, n.state.u_eq
        self._views["curr"] = self._views["next"]

    # ------------------------------------------------------------------ ghosts and neighbours
    def _set_ghost(self, side, cell, r, u):
        src = self._ghost_src.get(side)
        grad = any(rt.is_tensor(x) and x.requires_grad for x in (r, u))
        if src is not None and src[0] is r and src[1] is u and not grad:
            return                                        # same constants as last step: record still valid
        cell.state = ARZ.FullQ.from_r_u(r, u, self.speed_limit)
        self._ghost_src[side] = (r, u)

    def set_leftmost_cell(self, r, u):
        self._set_ghost(0, self.leftmost_cell, r, u)

    def set_rightmost_cell(self, r, u):
        self._set_ghost(1, self.rightmost_cell, r, u)

    def get_leftmost_cell(self):
        return self.leftmost_cell

    def get_rightmost_cell(self):
        return self.rightmost_cell

    def get_left_cell(self, id):
        return self.leftmost_cell if id == 0 else self.curr_cell[id - 1]

    def get_right_cell(self, id):
        return self.rightmost_cell if id == self.num_cell - 1 else self.curr_cell[id + 1]

    def which(self, pos):
        return math.floor(pos / self.cell_length)

    def add_flux_capacitor(self, next_lane_id, increment):
        self.flux_capacitor[next_lane_id] = self.flux_capacitor.get(next_lane_id, 0.0) + increment

    # ------------------------------------------------------------------ vector access (host tensors out)
    def _set_vector(self, which, rv, second, second_is_u):
        assert len(rv) == self.num_cell and len(second) == self.num_cell, "Cell number mismatch"
        r, s = rt.vector(rv), rt.vector(second)
        if second_is_u:
            e = ARZ.compute_u_eq(r, self.speed_limit)
            self._assign(which, r, r * (s - e), s, e)
        else:
            self._assign(which, r, s, ARZ.compute_u(r, s, self.speed_limit), ARZ.compute_u_eq(r, self.speed_limit))

    def set_state_vector_y(self, rv, yv):
        self._set_vector("curr", rv, yv, False)

    def set_state_vector_u(self, rv, uv):
        self._set_vector("curr", rv, uv, True)

    def set_next_state_vector_y(self, rv, yv):
        self._set_vector("next", rv, yv, False)

    def set_next_state_vector_u(self, rv, uv):
        self._set_vector("next", rv, uv, True)

    def _get_vector(self, which):
        rt.check_flags()
        return tuple(self._vec(which, k).cpu() for k in ("r", "y", "u"))

    def get_state_vector(self):
        """(density, relative flow, speed) as host tensors wired into the autograd graph."""
        return self._get_vector("curr")

    def get_next_state_vector(self):
        return self._get_vector("next")

    def device_state(self, which="curr"):
        """(r, y, u, u_eq) device vectors, no host copy."""
        return tuple(self._vec(which, k) for k in _FIELDS)

    def clear(self):
        for c in self.curr_cell + self.next_cell:
            c.state.clear()
        self._views = {"curr": None, "next": None}
        self.flux_capacitor.clear()
