"""Deferred stepping of the drop-in ``RoadNetwork``: consecutive ``forward`` calls become ONE fused rollout launch.

The reference's drivers advance a network one step at a time (``InverseProblem.simulate``,
example/inverse/_inverse.py:91-99) and only look at it again after the loop.  Stepping that loop call by call costs
one kernel launch per lane and step plus the autograd nodes around it; the kernels that take T steps in one launch
(``dhts_arz_rollout_*``, ``dhts_idm_rollout_*``, ``dhts_hyb_rollout_*``) already exist.  So ``RoadNetwork.forward``
only ENQUEUES a step, and the queue is executed -- as one rollout per lane group -- the first time anybody looks at
or changes something a step would have touched: ``lane.curr_cell``, ``lane.curr_vehicle``, ``get_state_vector``,
``set_state_vector*``, ``flux_capacitor``, ``network.vehicle`` / ``micro_route`` / ``macro_route`` ...
(properties that call ``flush`` first).  The device-side conditions (CFL, NaN gradient, collisions) are checked at
the flush instead of per step.

What is fused (anything else is stepped immediately, exactly as before):

  * ``iso``  every lane is disconnected (configs 1 and 2 of BASELINE.json: example/inverse/macro.py, micro.py).
    Macro lanes keep their own static ghost cells (road_network.py:312-321) -> ``ArzRolloutFn`` over all lanes of equal
    cell count; micro lanes follow the default ghost leader (road_network.py:441-443) -> ``IdmRolloutFn`` over all
    lanes as one CSR batch.  A head vehicle that would have left its lane during the queue (``micro_to_none``,
    conversion.py:202-215) is detected after the launch and the queue is then replayed step by step.
  * ``hyb``  connected networks of macro and micro lanes whose spawned vehicles are ``default_micro_vehicle``s
    (config 3: example/inverse/hybrid.py) -> ``hybrid_rollout`` in plain mode with R = 1; vehicles, routes, flux
    capacitors and ids are rebuilt from the kernel's state history, and ``np.random`` is advanced by replaying
    ``create_random_route`` for every spawn in the reference's order.

Only plain ``RoadNetwork`` objects are deferred: a subclass that overrides a boundary hook (ItscpRoadNetwork)
changes what a step means, so it keeps the immediate path.  Precision: in ``mixed`` mode the immediate path rounds
the state to fp32 after every step like the reference; a fused queue runs in fp64 and rounds once at the end
(differences are fp32 rounding, inside the fp32 tolerance of the parity tests).  ``DHTS_DEFER=0`` or
``runtime.configure(defer=False)`` switches the whole mechanism off.
"""
from __future__ import annotations

import numpy as np
import torch

from dhts_b200 import _lib
from dhts_b200 import functional as F
from dhts_b200.dropin import runtime as rt

_HOOKS = ("setup_boundary", "setup_macro_boundary", "setup_micro_boundary", "get_macro_boundary", "conversion",
          "conversion_macro", "conversion_micro", "get_macro_state_of_micro_lane")


def _plain_class(net) -> bool:
    """True when no step hook is overridden below the drop-in RoadNetwork (cached per class)."""
    cls = type(net)
    ok = cls.__dict__.get("_dhts_plain_hooks")
    if ok is None:
        from road.network.road_network import RoadNetwork
        ok = all(getattr(cls, h) is getattr(RoadNetwork, h) for h in _HOOKS)
        try:
            cls._dhts_plain_hooks = ok
        except Exception:
            pass
    return ok


def plan(net):
    """'iso' / 'hyb' / None for the network as it is now (lane set and links; cheap, recomputed on every enqueue)."""
    if not rt.defer_enabled() or not _plain_class(net):
        return None
    lanes = net._lane.values() if hasattr(net, "_lane") else net.lane.values()
    iso = True
    for lane in lanes:
        if lane.is_macro():
            if lane.record_case or lane.bdry_callback is not None:
                return None
        elif lane.record_flags:
            return None
        if lane.next_lane or lane.prev_lane:
            iso = False
    if iso:
        return "iso"
    return "hyb" if rt.defer_enabled("hyb") else None


# ------------------------------------------------------------------------------------------------ iso

def _macro_groups(net):
    groups = {}
    for lane in net.lane.values():
        if lane.is_macro():
            groups.setdefault(lane.num_cell, []).append(lane)
    return groups


def _flush_iso(net, steps, dt, replay):
    from model.macro._arz import ARZ
    sd, st = rt.step_dtype(), rt.store_dtype()
    dev, flags = rt.device(), rt.flags()
    results = []
    for N, lanes in _macro_groups(net).items():
        r = torch.stack([l._vec("curr", "r") for l in lanes]).to(sd)
        y = torch.stack([l._vec("curr", "y") for l in lanes]).to(sd)
        u = torch.stack([l._vec("curr", "u") for l in lanes]).to(sd).detach()
        ghost = torch.stack([torch.cat([l._ghost(c, k) for c in (l.leftmost_cell, l.rightmost_cell) for k in ("r", "y", "u")])
                             for l in lanes]).to(sd).reshape(len(lanes), 2, 3)
        key = ("geo", sd, dev)
        geo = net._defer_cache.get(key + tuple(id(l) for l in lanes))
        if geo is None:
            geo = (torch.tensor([float(l.cell_length) for l in lanes], dtype=sd, device=dev),
                   torch.tensor([float(l.speed_limit) for l in lanes], dtype=sd, device=dev))
            net._defer_cache[key + tuple(id(l) for l in lanes)] = geo
        ck = 1 if steps * len(lanes) * N * 16 <= (1 << 28) else 16
        try:
            rT, yT, uT = F.arz_rollout_state(r, y, u, ghost, geo[0], geo[1], float(dt), int(steps), ck, flags)
        except _lib.UnsupportedShape:
            return replay()
        results.append((lanes, rT, yT, uT))
    micro = [l for l in net.lane.values() if l.is_micro() and l._curr_vehicle]
    mres = None
    if micro:
        n_max = max(len(l._curr_vehicle) for l in micro)
        pv = [l._state() for l in micro]
        p = torch.cat([a for a, _ in pv]).to(sd); v = torch.cat([b for _, b in pv]).to(sd)
        par = torch.cat([l._params(sd)[0] for l in micro], dim=1).contiguous()
        counts = [len(l._curr_vehicle) for l in micro]
        off = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32, device=dev)
        from road.lane._micro_lane import DEFAULT_HEAD_POSITION_DELTA, DEFAULT_HEAD_SPEED_DELTA
        head = torch.tensor([[float(DEFAULT_HEAD_POSITION_DELTA), float(DEFAULT_HEAD_SPEED_DELTA)]], dtype=sd,
                            device=dev).repeat(len(micro), 1)
        try:
            pT, vT = F.idm_rollout_state(p, v, par, off, head, float(dt), int(steps), 16, flags, n_max)
        except _lib.UnsupportedShape:       # lanes longer than the fused kernel takes
            return replay()
        # micro_to_none (conversion.py:202-215) removes a head vehicle that reaches the end of its lane: the fused queue has
        # no list surgery, so such a queue is replayed step by step (positions never decrease: one test of the end state)
        ends = torch.tensor([float(l.length) for l in micro], dtype=sd, device=dev)
        heads = pT.detach()[(off[1:] - 1).long()]
        if bool((heads >= ends).any()):
            flags.reset()
            return replay()
        mres = (micro, counts, pT, vT)
    for lanes, rT, yT, uT in results:
        for i, l in enumerate(lanes):
            a, b, c = rT[i].to(st), yT[i].to(st), uT[i].to(st)
            e = ARZ.compute_u_eq(a, l.speed_limit)
            l._assign("curr", a, b, c, e)
            l._assign("next", a, b, c, e)
    if mres is not None:
        micro, counts, pT, vT = mres
        o = 0
        for l, n in zip(micro, counts):
            a, b = pT[o:o + n].to(st), vT[o:o + n].to(st)
            o += n
            l._hand_out(a, b)
            l._set_next(a, b)
            l.head_position_delta, l.head_speed_delta = DEFAULT_HEAD_POSITION_DELTA, DEFAULT_HEAD_SPEED_DELTA
    rt.check_flags()


# ------------------------------------------------------------------------------------------------ entry points

def flush(net):
    """Execute the queued steps of `net` (no-op when the queue is empty)."""
    steps = net._pending
    if not steps:
        return
    dt, diff, mode = net._pending_dt, net._pending_diff, net._pending_mode
    net._pending = 0                      # first: everything below reads lanes through the syncing properties

    def replay():
        for _ in range(steps):
            net._forward_now(dt, diff)

    if steps == 1 or mode is None:
        return replay()
    if mode == "iso":
        return _flush_iso(net, steps, dt, replay)
    from dhts_b200.dropin import deferred_hyb
    return deferred_hyb.flush_hyb(net, steps, dt, diff, replay)
