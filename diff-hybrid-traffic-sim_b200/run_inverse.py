"""``python -m dhts_b200.run_inverse --problem {macro,micro,hybrid}``: the reference's ``example/inverse/{macro,micro,
hybrid}.py`` (same arguments and defaults, same ``result/inverse/<run>/<method>/trial_<k>.txt`` files) with all trials of
the gradient-descent solver run as ONE batch on the fused rollouts.  ``--methods gd,nm,slsqp`` adds the scipy solvers of
``InverseProblem.evaluate`` (_inverse.py:99-172), trial by trial, over the batched objective."""
import argparse
import json
import time

import torch

from .inverse import HybridInverseBatch, MacroInverseBatch, MicroInverseBatch


def main(argv=None):
    parser = argparse.ArgumentParser("Script to solve inverse problems in traffic simulation")
    parser.add_argument("--problem", choices=["macro", "micro", "hybrid"], default="macro")
    parser.add_argument("--n_trial", type=int, default=5)
    parser.add_argument("--n_cell", type=int, default=10)
    parser.add_argument("--n_vehicle", type=int, default=10)
    parser.add_argument("--n_timestep", type=int, default=500)
    parser.add_argument("--cell_length", type=float, default=5.0)
    parser.add_argument("--vehicle_length", type=float, default=5.0)
    parser.add_argument("--speed_limit", type=float, default=30.0)
    parser.add_argument("--delta_time", type=float, default=0.01)
    parser.add_argument("--n_episode", type=int, default=100)
    parser.add_argument("--methods", type=str, default="gd")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--log_root", type=str, default="result/inverse")
    a = parser.parse_args(argv)
    if a.seed:
        torch.manual_seed(a.seed)
    run_name = "{}_{}".format(a.problem, time.time())
    common = (a.n_trial, a.n_timestep, a.n_episode, a.delta_time, a.speed_limit, run_name)
    if a.problem == "micro":
        prob = MicroInverseBatch(*common, a.n_vehicle, a.vehicle_length, log_root=a.log_root)
    else:
        cls = MacroInverseBatch if a.problem == "macro" else HybridInverseBatch
        prob = cls(*common, a.n_cell, a.cell_length, log_root=a.log_root)
    t0 = time.time()
    est = prob.initialize()
    summary = {"problem": a.problem, "run": prob.log_dir, "trials": a.n_trial, "episodes": a.n_episode, "methods": {}}
    for m in a.methods.split(","):
        t1 = time.time()
        if m == "gd":
            beg, end = prob.solve_gd(est)
        else:
            beg, end = [], []
            per = []
            for k in range(a.n_trial):
                row = tuple(s[k] for s in est)
                per.append(prob.solve_scipy(row, {"nm": "Nelder-Mead", "slsqp": "SLSQP"}[m], k) if m in ("nm", "slsqp")
                           else prob.solve_cma(row, 1.0, k))
            beg = [[per[k][0][e] for k in range(a.n_trial)] for e in range(a.n_episode)]
            end = [[per[k][1][e] for k in range(a.n_trial)] for e in range(a.n_episode)]
        torch.cuda.synchronize()
        prob.write_trials(m, beg, end)
        summary["methods"][m] = {"wall_s": round(time.time() - t1, 3), "end_error_first": end[0], "end_error_last": end[-1]}
    summary["wall_s"] = round(time.time() - t0, 3)
    print(json.dumps(summary))
    return summary


if __name__ == "__main__":
    main()
