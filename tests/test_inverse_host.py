"""Host side of dhts_b200.inverse (no GPU): draws, bounds, vector packing, trial_<k>.txt format, no CPU fallback."""
import numpy as np
import pytest
import torch


def test_draws_bounds_and_formats(tmp_path):
    from dhts_b200.inverse import MacroInverseBatch, MicroInverseBatch, log_error
    prob = MacroInverseBatch(3, 50, 2, 0.01, 30.0, "run", 10, 5.0, device="cpu", log_root=str(tmp_path))
    torch.manual_seed(1)
    prob._trial_beg = None
    prob.init_network(0)
    b = prob.random_initial_state(0)
    assert b[0].dtype == torch.float32 and b[0].shape == (10,) and float(b[1].max()) <= 30.0
    # the reference's draw order per trial: bdry density(2), bdry speed(2), state(10+10) [discarded], state(10+10)  (macro.py:35-66,
    # _inverse.py:80-84)
    torch.manual_seed(1)
    bd = torch.rand(2); bs = torch.rand(2) * 30.0; torch.rand(10); torch.rand(10)
    r = torch.rand(10); u = torch.rand(10) * torch.tensor([30.0])
    assert torch.equal(prob._bd[0], bd) and torch.equal(prob._bs[0], bs) and torch.equal(b[0], r) and torch.equal(b[1], u)
    prob._trial_beg = b
    e = prob.random_initial_state(0)
    assert float((e[0] - b[0]).abs().max()) < 0.06 and float(e[0].min()) >= 0.0 and float(e[0].max()) <= 1.0
    lb, ub = prob.bounds()
    assert float(lb[0].max()) == 0 and float(ub[0].min()) == 1 and float(ub[1].min()) == 30
    v = prob.vectorize((torch.zeros(3, 10), torch.ones(3, 10)))
    a, c = prob.unvectorize(v)
    assert v.shape == (3, 20) and float(a.sum()) == 0 and float(c.sum()) == 30
    m = MicroInverseBatch(2, 50, 2, 0.01, 30.0, "run", 10, 5.0, device="cpu", log_root=str(tmp_path))
    m._trial_beg = None
    p, s = m.random_initial_state(0)
    lbm, ubm = m.bounds()
    assert bool((p >= lbm[0].float()).all()) and bool((p <= ubm[0].float()).all()) and 9.0 <= float(s.min()) and float(s.max()) <= 21.0
    log_error(str(tmp_path / "trial_0.txt"), [1.5, 0.25], [3.0, 2.0])
    assert open(str(tmp_path / "trial_0.txt")).read() == "1.5 3.0\n0.25 2.0\n"       # _inverse.py:504-514
    paths = prob.write_trials("cma", [[1, 2, 3], [4, 5, 6]], [[7, 8, 9], [1, 1, 1]])
    assert paths[1].endswith("run/cma-es/trial_1.txt") and open(paths[1]).read() == "2 8\n5 1\n"


def test_no_cpu_fallback(tmp_path):
    from dhts_b200.inverse import MacroInverseBatch
    prob = MacroInverseBatch(1, 5, 1, 0.01, 30.0, "run", 10, 5.0, device="cpu", log_root=str(tmp_path))
    prob._bd = [torch.rand(2)]; prob._bs = [torch.rand(2)]
    prob._finish_networks()
    with pytest.raises(Exception):
        prob.simulate((torch.rand(1, 10, dtype=torch.float64), torch.rand(1, 10, dtype=torch.float64)), False)
