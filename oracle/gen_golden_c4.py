"""TEST INFRASTRUCTURE: freeze one FULL config-4 episode of the live reference (run_itscp_hybrid.sh: hybrid mode,
3x3 intersections, 1 lane per road, access lanes of 5 m, u_max 60, 20 s policy = 600 frames at 30 Hz, 4 s signals =
45 actions) into tests/golden/itscp_c4_fp64.npz.

Run in THIS container only (needs /root/reference; several minutes):

    python oracle/gen_golden_c4.py

The episode is what ``Trainer.run_episode(True)`` makes ``ItscpEnv.step`` do (example/control/trainer.py:160-196,
example/control/itscp/_env.py:537-742): lanes start empty, boundary lanes receive the ``problem_1`` schedule
(example/control/itscp/problem.py:5-68, restated below with the same ``np.random`` draws), signals come from the
action through ``lane_signal_info``, the reward is the queue-length term with the running-mean sigmoid constants.
highway-env / gym / pygame are absent here, so ``_env.py`` itself cannot be imported: ``gen_golden_hyb.run_case``
restates its driving loop around the UNMODIFIED ``ItscpRoadNetwork`` + lanes.  Only inputs, the reward, its action
gradient and thinned state snapshots are stored (the per-sample sigmoid constants are NOT: the headless env has to
reproduce them).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import gen_golden_hyb as H  # noqa: E402


def problem_schedule(num_session):
    """example/control/itscp/problem.py:5-68 over lane infos; returns [T, L]."""
    def cb(lanes, T):
        per = T // num_session
        dirs = []
        for i in range(num_session):
            if i == 0:
                dirs.append("NS" if np.random.random((1)).item() > 0.5 else "WE")
            else:
                dirs.append("WE" if dirs[-1] == "NS" else "NS")
        out = np.zeros((T, len(lanes)))
        for l, info in enumerate(lanes):
            cur = []
            for s in range(num_session):
                r = np.random.random((1)).item()
                ns = info.loc in ("north", "south")
                we = info.loc in ("west", "east")
                hot = ns if dirs[s] == "NS" else we
                r = 0.9 + r * 0.1 if hot else 0.0 + r * 0.01
                cur.extend([r] * per)
            cur = cur[:T]
            out[:len(cur), l] = cur
        return out
    return cb


def main():
    H.switch_fp64()
    from dhts_b200.itscp import ItscpGrid
    T = int(os.environ.get("C4_T", "600"))
    H.CAP = 8
    out = H.run_case("c", ItscpGrid(3, 1, 5.0, 5.0), T=T, frames_per_signal=120, seed=int(os.environ.get("C4_SEED", "4")),
                     empty=True, schedule=problem_schedule(1), action_range=(0.3, 0.7))
    keep = {}
    every = 50
    for k, v in out.items():
        name = k[2:]
        if name in ("kcell", "kveh", "sig", "g_sig", "g_inc", "head", "vid", "w_veh"):
            continue
        if name in ("hist", "veh"):
            keep[k + "_thin"] = v[::every]
            continue
        keep[k] = v
    keep["c_every"] = every
    keep["c_w_veh"] = out["c_w_veh"]
    path = os.path.join(H.OUT, "itscp_c4_fp64.npz")
    np.savez_compressed(path, **keep)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
