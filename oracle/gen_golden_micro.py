"""TEST INFRASTRUCTURE: freeze episodes of the LIVE reference's ITSCP network in MICRO mode into tests/golden/.

Run in THIS container only (needs /root/reference):

    python oracle/gen_golden_micro.py

``run_itscp_micro.sh`` (mode=micro, 1 intersection x 3 lanes, 30 m access lanes, u_max 60, 10 s policy = 300 frames
at 30 Hz, 2 s signals = 5 actions): every lane is a plain ``MicroLane`` (_env.py:484-488), boundary lanes are fed
from the stochastic waiting list of ``ItscpRoadNetwork.setup_micro_boundary`` (_simulator.py:153-174: one
``np.random.random`` draw per boundary lane and frame WHILE the lane has room, a vehicle enters when the draw is
below the scheduled inflow), head deltas are signal-blended (_simulator.py:176-276) and the gradient reaches the
action by plain autograd through ``MicroLane.forward`` (_micro_lane.py:131-186).

As in gen_golden_hyb.py the reference's ``ItscpRoadNetwork`` and lanes are imported unmodified and the pieces of
``_env.py`` that drive them (``_make_micro_route`` :202-219, ``_simulate_step`` :590-742, ``_reward`` :770-797) are
restated here, because ``_env.py`` itself needs highway-env / gym / pygame.

Recorded INPUTS: action, inflow schedule, the waiting routes of every lane (in list order; the reference pops from
the END) and the uniform draws the episode consumed, in order.  Recorded OUTPUTS: every frame's vehicles (head
first), counts, head deltas, the reward and its action gradient.

Fixture: tests/golden/itscp_micro_fp64.npz.
"""
import os
import sys

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DHTS_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.path.insert(2, HERE)

from gen_golden import switch_fp64  # noqa: E402
from gen_golden_net import load_reference_simulator, f  # noqa: E402
from gen_golden_c4 import problem_schedule  # noqa: E402

RLEN = 32        # MAX_ROUTE_LENGTH, road_network.py:15


def run_case(tag, grid, T, frames_per_signal, seed, K=10, cap=8, umax=60.0, freq=30, static_speed=0.2, schedule=None,
             action_range=(0.1, 0.9)):
    from dmath.operation import sigmoid
    from road.lane._micro_lane import MicroLane
    ItscpRoadNetwork, RunningMean = load_reference_simulator()

    class _NP64:
        float32 = np.float64

        def __getattr__(self, k):
            return getattr(np, k)
    sys.modules["example.common.rms"].np = _NP64()      # RunningMean accumulates in the state dtype
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    dt = 1.0 / freq
    n = grid.num_intersection
    net = ItscpRoadNetwork(umax)
    for info in grid.lanes:
        net.add_lane(MicroLane(len(net.lane), info.length, umax))          # _env.py:484-488
    for a, b in grid.links:
        net.connect_lane(a, b)
    L = grid.L
    n2 = n * n
    n_phase = max(1, T // frames_per_signal)
    action = th.tensor(rng.uniform(action_range[0], action_range[1], n_phase * n2), requires_grad=True)
    bl = grid.boundary_lanes()
    inc_np = np.zeros((T, L))
    full = schedule(grid.lanes, T)                      # the schedule callback draws from np.random first, as env.reset does
    for l in bl:
        inc_np[:, l] = full[:, l]
    for _ in range(T):                                  # _make_macro_route: draws even though no lane is macro
        net.create_random_macro_route()
    wait = -np.ones((L, K, RLEN), dtype=np.int32)       # _make_micro_route, _env.py:202-219: every lane, K vehicles
    for l in range(L):
        net.lane_waiting_micro_vehicle[l] = []
        net.lane_waiting_micro_route[l] = []
        for k in range(K):
            nv, nr = net.create_default_vehicle_with_random_route(l)
            net.lane_waiting_micro_vehicle[l].append(nv)
            net.lane_waiting_micro_route[l].append(nr)
            wait[l, k, :len(nr.route)] = nr.route
    state_after_reset = np.random.get_state()
    draws = []
    real_random = np.random.random

    def recording_random(*a, **k):
        x = real_random(*a, **k)
        draws.append(float(np.asarray(x).reshape(-1)[0]))
        return x
    np.random.random = recording_random
    rms = RunningMean(100_000)
    veh = np.zeros((T + 1, L, cap, 3)); vcnt = np.zeros((T + 1, L), dtype=np.int32)
    head = np.zeros((T, L, 2)); kveh = np.zeros((T, L, cap))
    reward = 0

    def snapshot(t):
        for l in range(L):
            lane = net.lane[l]
            assert lane.num_vehicle() <= cap, "raise cap"
            vcnt[t, l] = lane.num_vehicle()
            for j, mv in enumerate(reversed(lane.curr_vehicle)):      # front (head) first
                veh[t, l, j] = (f(mv.position), f(mv.speed), f(mv.a))

    snapshot(0)
    try:
        for t in range(T):
            phase = min(t // frames_per_signal, n_phase - 1)
            progress = min((t % frames_per_signal) / frames_per_signal, 1.0)
            for l, info in enumerate(grid.lanes):
                if info.loc == "mid" or not info.approaching:
                    s = 1.0
                else:
                    a = action[phase * n2 + info.row * n + info.col]
                    s = sigmoid(a - progress, constant=32) if info.loc in ("west", "east") else sigmoid(progress - a, constant=32)
                net.lane_signal[l] = s
                net.lane_incoming[l] = inc_np[t, l] if l in bl else -1
            net.forward(dt, True)
            for l in range(L):
                head[t, l] = (f(net.lane[l].head_position_delta), f(net.lane[l].head_speed_delta))
            snapshot(t + 1)
            for l in range(L):                                            # _env.py:662-742, lanes in id order, tail first
                lane = net.lane[l]
                q = 0
                for j, mv in enumerate(lane.curr_vehicle):
                    speed = mv.speed if isinstance(mv.speed, th.Tensor) else th.tensor(mv.speed)
                    with th.no_grad():
                        rms.update((static_speed - speed).cpu().numpy())
                        constant = 16.0 / np.abs(rms.mean())
                    kveh[t, l, lane.num_vehicle() - 1 - j] = constant
                    q = q + sigmoid(static_speed - speed, constant=constant)
                reward = reward + (-1.0) * ((q ** 2.0) * dt)
    finally:
        np.random.random = real_random
    reward.backward()
    rng_next = float(np.random.random())               # where the episode left np.random
    out = dict(T=T, frames_per_signal=frames_per_signal, umax=umax, dt=dt, veh_len=net.vehicle_length, static_speed=static_speed,
               num_intersection=n, num_lane=grid.num_lane, lane_length=grid.lane_length, K=K, cap=cap, seed=seed,
               action=action.detach().numpy(), incoming=inc_np, wait=wait, draws=np.array(draws), veh=veh, vcnt=vcnt,
               head=head, kveh=kveh, reward=f(reward), g_action=action.grad.numpy(), rng_next=rng_next,
               rng_after_reset_pos=np.int64(state_after_reset[2]), rng_after_reset_key=state_after_reset[1])
    print(tag, "L", L, "T", T, "draws", len(draws), "vehicles entered", net.num_vehicle, "max on lane", vcnt.max(),
          "reward %.6g |g_action| %.3g" % (out["reward"], np.abs(out["g_action"]).max()))
    return {tag + "_" + k: v for k, v in out.items()}


def main():
    switch_fp64()
    from dhts_b200.itscp import ItscpGrid
    out = {}
    # run_itscp_micro.sh's geometry and lengths; problem_1 schedule
    out.update(run_case("m", ItscpGrid(1, 3, 30.0, 5.0), T=int(os.environ.get("MIC_T", "300")), frames_per_signal=60, seed=31,
                        schedule=problem_schedule(1), action_range=(0.3, 0.7)))
    # two intersections per side: hand-offs between intersections, longer routes, problem_2 schedule
    out.update(run_case("n", ItscpGrid(2, 1, 20.0, 5.0), T=int(os.environ.get("MIC_T2", "240")), frames_per_signal=60, seed=32,
                        schedule=problem_schedule(2), action_range=(0.2, 0.8)))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "itscp_micro_fp64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
