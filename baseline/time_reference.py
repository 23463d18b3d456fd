"""Time the reference's OWN CPU path (the unmodified Python lanes of baseline/_ref) on this box's host cores, on
lanes of the benchmark's shape (SURVEY 8d "CPU baseline"):

  * ``dMacroLane`` with N = 1024 cells, dx = 5, u_max = 30, dt = 0.01: 20 steps of ``RoadNetwork.forward`` +
    one ``loss.backward()`` through them (road/lane/dmacro_lane.py:68-85,234-310);
  * ``dMicroLane`` with n = 64 vehicles (per-vehicle parameters from the ``random_micro_vehicle`` ranges,
    road/vehicle/micro_vehicle.py:88-109): 100 steps + backward (road/lane/dmicro_lane.py:63-77,228-297).

The reference is single-threaded Python, so ONE PROCESS PER HOST CORE runs an independent lane; the aggregate rate
is (processes x updates) / (the slowest process's time).  Called by bench.py (cpu_baseline.reference_python and
``--impl reference``); prints one JSON line when run as a script.

    python baseline/time_reference.py [--procs P] [--macro-steps 20] [--micro-steps 100]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def worker(a):
    import numpy as np
    import torch as th
    th.set_num_threads(1)
    from baseline import install_ref
    install_ref.add_to_path(with_core=True)
    from road.lane.dmacro_lane import dMacroLane
    from road.lane.dmicro_lane import dMicroLane
    from road.network.road_network import RoadNetwork
    from road.vehicle.micro_vehicle import MicroVehicle
    import road.lane.dmacro_lane as probe
    rng = np.random.default_rng(20221008 + a.worker)
    N, dx, umax, dt = a.cells, 5.0, 30.0, 0.01
    out = {"core_packages_from": os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(probe.__file__))))}

    # ---- macro lane (example/inverse/macro.py:36-100 builds the same one-lane network)
    tr = th.tensor(rng.uniform(0, 1, N), dtype=th.float32, requires_grad=True)
    tu = th.tensor(rng.uniform(0, 1, N) * umax, dtype=th.float32, requires_grad=True)
    lane = dMacroLane(0, N * dx, umax, dx)
    lane.set_state_vector_u(tr, tu)
    lane.set_leftmost_cell(float(rng.uniform()), float(rng.uniform() * umax))
    lane.set_rightmost_cell(float(rng.uniform()), float(rng.uniform() * umax))
    net = RoadNetwork(umax); net.add_lane(lane)
    t0 = time.perf_counter()
    for _ in range(a.macro_steps):
        net.forward(dt, True)
    r, y, u = lane.get_state_vector()
    t1 = time.perf_counter()
    ((r - 0.5) ** 2).sum().add(((u - 15.0) ** 2).sum()).backward()
    t2 = time.perf_counter()
    out.update(macro_fwd_s=t1 - t0, macro_bwd_s=t2 - t1, macro_updates=N * a.macro_steps,
               macro_grad_finite=bool(th.isfinite(tr.grad).all()))

    # ---- micro lane (example/inverse/micro.py:36-58)
    n = a.vehicles
    p = th.tensor(np.arange(n) * 20.0 + rng.uniform(0, 10, n), dtype=th.float32, requires_grad=True)
    v = th.tensor(rng.uniform(9, 21, n), dtype=th.float32, requires_grad=True)
    mlane = dMicroLane(0, 1e10, umax)
    for i in range(n):
        mlane.add_head_vehicle(MicroVehicle(i, p[i], v[i], float(rng.uniform(1.5, 2.0) * umax), float(rng.uniform(1.0, 1.5) * umax),
                                            float(rng.uniform(0.8, 1.2) * umax), float(rng.uniform(1, 2)),
                                            float(rng.uniform(0.2, 0.6)), 5.0, 5.0))
    t0 = time.perf_counter()
    for _ in range(a.micro_steps):
        mlane.forward(dt); mlane.update_state()
    pT, vT = mlane.get_state_vector()
    t1 = time.perf_counter()
    ((pT - p.detach() - 15.0) ** 2).sum().add(((vT - 15.0) ** 2).sum()).backward()
    t2 = time.perf_counter()
    out.update(micro_fwd_s=t1 - t0, micro_bwd_s=t2 - t1, micro_updates=n * a.micro_steps,
               micro_grad_finite=bool(th.isfinite(p.grad).all()))
    print(json.dumps(out), flush=True)


def measure(procs=None, macro_steps=20, micro_steps=100, cells=1024, vehicles=64):
    """One worker process per host core; returns the aggregate rates and the sample description."""
    from baseline import install_ref
    if not install_ref.available():
        return None
    procs = procs or len(os.sched_getaffinity(0))
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(i), "--macro-steps", str(macro_steps),
                            "--micro-steps", str(micro_steps), "--cells", str(cells), "--vehicles", str(vehicles)],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for i in range(procs)]
    rows = []
    for p in ps:
        o, e = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference worker failed: " + e[-2000:])
        rows.append(json.loads(o.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    tm = max(r["macro_fwd_s"] + r["macro_bwd_s"] for r in rows)
    ti = max(r["micro_fwd_s"] + r["micro_bwd_s"] for r in rows)
    return {"value": procs * rows[0]["macro_updates"] / tm, "unit": "cell-updates/s", "cores": procs,
            "idm_value": procs * rows[0]["micro_updates"] / ti, "idm_unit": "vehicle-updates/s",
            "per_core": rows[0]["macro_updates"] / tm, "idm_per_core": rows[0]["micro_updates"] / ti,
            "sample": "%d processes (one per host core), each one live dMacroLane(N=%d) x %d steps fwd+bwd and one "
                      "dMicroLane(n=%d) x %d steps fwd+bwd; rate = all processes' updates / slowest process" %
                      (procs, cells, macro_steps, vehicles, micro_steps),
            "kind": "reference", "code": rows[0]["core_packages_from"], "wall_s": wall,
            "grads_finite": all(r["macro_grad_finite"] and r["micro_grad_finite"] for r in rows)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", type=int, default=-1)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--macro-steps", type=int, default=20)
    ap.add_argument("--micro-steps", type=int, default=100)
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--vehicles", type=int, default=64)
    a = ap.parse_args()
    if a.worker >= 0:
        worker(a)
    else:
        print(json.dumps(measure(a.procs or None, a.macro_steps, a.micro_steps, a.cells, a.vehicles)))
