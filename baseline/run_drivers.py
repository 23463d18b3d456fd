"""Run the reference's UNMODIFIED inverse-problem drivers (configs 1-3 of BASELINE.json) for a few gradient-descent
episodes, either on the reference's own CPU lanes or on top of the B200 drop-in packages.

    python baseline/run_drivers.py --impl reference --problem macro --episodes 3 --out x.npz
    python baseline/run_drivers.py --impl dropin    --problem hybrid --episodes 3 --out y.npz   # needs a GPU

The driver classes come from ``baseline/_ref/example/inverse/{macro,micro,hybrid}.py`` (file-for-file copies of
the reference, see baseline/install_ref.py) in BOTH arms; what differs is which ``road/ model/ dmath/`` packages
they import: the reference's (``--impl reference``) or ``dhts_b200.dropin`` (``--impl dropin``).  The calls
below are the body of ``InverseProblem.evaluate`` for the gradient-descent method (example/inverse/_inverse.py:
113-134 -> solve_gd :185-242) and of each script's ``__main__`` (macro.py:243-269, micro.py:238-265,
hybrid.py:256-282), with the defaults of those scripts.

Prints one JSON line (seconds per episode etc.); --out stores the error curves and the final estimate.
Run as a separate process per (impl, problem): the two arms import different modules under the same names.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["reference", "dropin"], required=True)
    ap.add_argument("--problem", choices=["macro", "micro", "hybrid"], required=True)
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--timesteps", type=int, default=500)
    ap.add_argument("--seed", type=int, default=20221008)
    ap.add_argument("--precision", default="mixed", help="drop-in arithmetic: mixed (= the reference as shipped) / float64 / float32")
    ap.add_argument("--cpu-standin", action="store_true",
                    help="TEST ONLY: route the drop-in's kernel calls through tests/cpu_standin.py (host-logic check on a box without a GPU)")
    ap.add_argument("--fp64", action="store_true",
                    help="run everything in float64: the reference through the no-edit dtype rebinding of SURVEY App. C "
                         "(its modules' np.float32 / th.float32 names point at the 64-bit types), the drop-in with "
                         "precision=float64; the drivers' own tensors via problem.dtype and torch's default dtype")
    ap.add_argument("--no-defer", action="store_true", help="drop-in: step every RoadNetwork.forward immediately")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.fp64:
        a.precision = "float64"

    import numpy as np
    import torch as th
    from baseline import install_ref
    th.set_num_threads(1)
    ctx = None
    if a.impl == "dropin":
        import dhts_b200.dropin as dropin
        dropin.install(precision=a.precision)
        if a.no_defer:
            dropin.runtime.configure(defer=False)
        install_ref.add_to_path(with_core=False)
        if a.cpu_standin:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import cpu_standin
            ctx = cpu_standin.patched(precision=a.precision)
            ctx.__enter__()
    else:
        install_ref.add_to_path(with_core=True)
    if a.fp64:
        th.set_default_dtype(th.float64)
        if a.impl == "reference":
            class _NP64:
                float32 = np.float64

                def __getattr__(self, k):
                    return getattr(np, k)

            class _TH64:
                float32 = th.float64

                def __getattr__(self, k):
                    return getattr(th, k)

            import model.macro.darz as m0, road.lane.dmacro_lane as m1, road.lane._macro_lane as m2
            import road.lane.dmicro_lane as m3, road.lane._micro_lane as m4, model.micro.didm as m5
            for m in (m0, m1, m2, m3):
                m.np = _NP64()
            for m in (m1, m2, m3, m4, m5):
                m.th = _TH64()
    import road.lane.dmacro_lane as probe
    core = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(probe.__file__))))

    os.chdir(os.environ.get("DHTS_RUN_DIR", "/tmp"))        # the drivers create result/inverse/<run_name> in the cwd
    n, dt, umax, T = 10, 0.01, 30.0, a.timesteps
    run_name = "%s_%s_%d" % (a.problem, a.impl, os.getpid())
    if a.problem == "macro":
        from example.inverse.macro import MacroInverseProblem
        problem = MacroInverseProblem(1, T, a.episodes, dt, umax, run_name, n, 5.0)
    elif a.problem == "micro":
        import example.inverse.micro as micro_mod
        micro_mod.speed_limit = umax            # micro.py:111 reads the module global its __main__ sets (:252)
        problem = micro_mod.MicroInverseProblem(1, T, a.episodes, dt, umax, run_name, n, 5.0)
        problem.gd_lr = 1e-2                    # micro.py:264
    else:
        from example.inverse.hybrid import HybridInverseProblem
        problem = HybridInverseProblem(1, T, a.episodes, dt, umax, run_name, n, 5.0)

    if a.fp64:
        problem.dtype = th.float64
    # torch imports its compiler stack the first time an optimizer is built and stepped (seconds, once per
    # process): do that before the clock starts
    w = th.zeros(2, requires_grad=True)
    o = th.optim.Adam([w], lr=1e-3); w.sum().backward(); o.step()
    th.manual_seed(a.seed); np.random.seed(a.seed % (2 ** 31))
    sync = (lambda: th.cuda.synchronize()) if (a.impl == "dropin" and th.cuda.is_available()) else (lambda: None)
    t0 = time.perf_counter()
    problem.initialize()                        # _inverse.py:66-88 (one non-differentiable rollout)
    est = problem.random_initial_state()
    sync(); t1 = time.perf_counter()
    beg, end = problem.solve_gd(est, problem.gd_lr)
    sync(); t2 = time.perf_counter()
    if ctx is not None:
        ctx.__exit__(None, None, None)
    line = {"impl": a.impl, "problem": a.problem, "episodes": a.episodes, "timesteps": T,
            "s_per_episode": (t2 - t1) / max(1, a.episodes), "init_s": t1 - t0, "end_errors": end, "beg_errors": beg,
            "core_packages_from": core, "drivers_from": os.path.dirname(os.path.abspath(sys.modules[type(problem).__module__].__file__))}
    if a.out:
        np.savez(a.out, end_errors=np.asarray(end), beg_errors=np.asarray(beg),
                 est0=np.stack([np.asarray(e.detach().cpu(), dtype=np.float64) for e in est]),
                 end_state=np.stack([np.asarray(e.detach().cpu(), dtype=np.float64) for e in problem.end_state]),
                 beg_state=np.stack([np.asarray(e.detach().cpu(), dtype=np.float64) for e in problem.beg_state]))
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
