"""GPU: population-parallel inverse problems (dhts_b200.inverse) against the Adam curves of the live reference's three
inverse problems (tests/golden/inverse_fp64.npz, frozen by oracle/gen_golden.py from example/inverse/{macro,micro,hybrid}
loops in fp64).  Every trial row of a batch must reproduce the single-trial reference curve; fp64 tolerances at the asserts."""
import numpy as np
import pytest
import torch

from conftest import golden, relerr

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _t(a, dev, P):
    return torch.tensor(np.asarray(a, dtype=np.float64), dtype=F64, device=dev).unsqueeze(0).expand(P, -1).contiguous()


@pytest.mark.parametrize("mode", ["macro", "hybrid"])
def test_macro_and_hybrid_batches_follow_the_reference_curves(dev, mode, tmp_path):
    from dhts_b200.inverse import HybridInverseBatch, MacroInverseBatch
    g = golden("inverse_fp64")
    N, dx, umax, dt, T, E = int(g["N"]), float(g["dx"]), float(g["umax"]), float(g["dt"]), int(g["T"]), int(g["episodes"])
    P = 3
    cls = MacroInverseBatch if mode == "macro" else HybridInverseBatch
    prob = cls(P, T, E, dt, umax, "run", N, dx, device=dev, log_root=str(tmp_path))
    prob._bd = [torch.tensor(g[mode + "_bd"])] * P; prob._bs = [torch.tensor(g[mode + "_bs"])] * P
    prob._finish_networks()
    prob.beg_state = (_t(g[mode + "_true_r"], dev, P), _t(g[mode + "_true_u"], dev, P))
    with torch.no_grad():
        prob.end_state = prob.simulate(prob.beg_state, False)
    prob.flags.check()
    assert relerr(prob.end_state[0][1].cpu(), g[mode + "_end_r"]) < 1e-9 and relerr(prob.end_state[1][1].cpu(), g[mode + "_end_u"]) < 1e-9
    est = (_t(g[mode + "_est_r"], dev, P), _t(g[mode + "_est_u"], dev, P))
    beg, end = prob.solve_gd(est, 1e-3)
    end = np.array(end)                                                   # [E, P]
    for p in range(P):
        assert relerr(end[:, p], g[mode + "_errs"]) < 1e-8
        assert relerr(prob.estimate[0][p].cpu(), g[mode + "_final_r"]) < 1e-8
        assert relerr(prob.estimate[1][p].cpu(), g[mode + "_final_u"]) < 1e-8
    # population evaluation == the first GD episode's error, candidate by candidate, on any trial's network
    v = prob.vectorize(est)
    b, e = prob.evaluate_vector_states(torch.cat([v, v]), torch.tensor([0, 1, 2, 2, 1, 0]))
    assert relerr(e.cpu(), np.full(6, g[mode + "_errs"][0])) < 1e-8 and relerr(b.cpu(), np.array(beg[0] * 2)) < 1e-12
    paths = prob.write_trials("gd", beg, end.tolist())
    rows = [l.split() for l in open(paths[2]).read().splitlines()]
    assert len(paths) == P and len(rows) == E and abs(float(rows[0][1]) - g[mode + "_errs"][0]) < 1e-8 * g[mode + "_errs"][0]


def test_micro_batch_follows_the_reference_curve(dev, tmp_path):
    from dhts_b200.inverse import MicroInverseBatch
    g = golden("inverse_fp64")
    n, umax, dt, T, E = int(g["n"]), float(g["umax"]), float(g["dt"]), int(g["T"]), int(g["episodes"])
    P = 4
    prob = MicroInverseBatch(P, T, E, dt, umax, "run", n, 5.0, device=dev, log_root=str(tmp_path))
    prob.beg_state = (_t(g["micro_true_p"], dev, P), _t(g["micro_true_v"], dev, P))
    with torch.no_grad():
        prob.end_state = prob.simulate(prob.beg_state, False)
    assert relerr(prob.end_state[0][2].cpu(), g["micro_end_p"]) < 1e-9 and relerr(prob.end_state[1][2].cpu(), g["micro_end_v"]) < 1e-9
    est = (_t(g["micro_est_p"], dev, P), _t(g["micro_est_v"], dev, P))
    # the reference loop of the fixture (oracle/gen_golden.py) applies no bound projection: neutralise it
    prob.bounds = lambda: ((est[0][0] * 0 - 1e30, est[0][0] * 0 - 1e30), (est[0][0] * 0 + 1e30, est[0][0] * 0 + 1e30))
    beg, end = prob.solve_gd(est, 1e-2)
    end = np.array(end)
    for p in range(P):
        assert relerr(end[:, p], g["micro_errs"]) < 1e-8
        assert relerr(prob.estimate[0][p].cpu(), g["micro_final_p"]) < 1e-8
        assert relerr(prob.estimate[1][p].cpu(), g["micro_final_v"]) < 1e-8


def test_seeded_initialize_and_trials_are_independent(dev, tmp_path):
    """initialize() draws per trial in the reference's order; a 5-trial batch gives the same curve for trial k as a
    batch that holds only trials 0..k (rows never interact)."""
    from dhts_b200.inverse import MacroInverseBatch
    curves = []
    for P in (5, 2):
        torch.manual_seed(20221008)
        prob = MacroInverseBatch(P, 100, 4, 0.01, 30.0, "r%d" % P, 10, 5.0, device=dev, log_root=str(tmp_path))
        prob.initialize()
        beg, end = prob.solve_gd()
        curves.append(np.array(end))
        assert prob.beg_state[0].shape == (P, 10) and float(prob.beg_state[0].min()) >= 0 and float(prob.beg_state[1].max()) <= 30
        d = (prob.initial_estimate[0] - prob.beg_state[0]).abs().max()
        assert 0 < float(d) < 0.06                                             # estimate = truth + N(0, 1e-2), clamped
    assert np.array_equal(curves[0][:, :2], curves[1])
    assert (curves[0][-1] < curves[0][0]).all()                               # Adam reduces every trial's error


def test_scipy_solvers_over_the_batched_objective(dev, tmp_path):
    """Nelder-Mead and SLSQP (_inverse.py:301-352) over evaluate_vector_states: error lists of num_episode entries, the
    first entry is the error of the initial estimate, SLSQP (batched finite-difference gradient) reduces the error."""
    from dhts_b200.inverse import MacroInverseBatch
    torch.manual_seed(5)
    prob = MacroInverseBatch(2, 100, 12, 0.01, 30.0, "s", 10, 5.0, device=dev, log_root=str(tmp_path))
    est = prob.initialize()
    row = tuple(s[1] for s in est)
    b0, e0 = prob.evaluate_vector_states(prob.vectorize(row)[None], 1)
    for method in ("Nelder-Mead", "SLSQP"):
        beg, end = prob.solve_scipy(row, method, 1)
        assert len(beg) == len(end) == 12 and abs(end[0] - float(e0[0])) < 1e-12 and abs(beg[0] - float(b0[0])) < 1e-12
        assert min(end) <= end[0]
        if method == "SLSQP":
            assert min(end) < 0.9 * end[0]


def test_run_inverse_cli_writes_the_reference_tree(dev, tmp_path):
    import os
    from dhts_b200.run_inverse import main
    for problem in ("macro", "micro", "hybrid"):
        out = main(["--problem", problem, "--n_trial", "3", "--n_timestep", "120", "--n_episode", "5", "--seed", "3",
                    "--log_root", str(tmp_path)])
        d = os.path.join(out["run"], "gd")
        assert sorted(os.listdir(d)) == ["trial_0.txt", "trial_1.txt", "trial_2.txt"]
        rows = [l.split() for l in open(os.path.join(d, "trial_1.txt"))]
        assert len(rows) == 5 and all(len(r) == 2 for r in rows)
        last, first = out["methods"]["gd"]["end_error_last"], out["methods"]["gd"]["end_error_first"]
        assert all(np.isfinite(last)) and sum(last) < sum(first)
