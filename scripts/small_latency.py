"""Kernel latency of the reference-sized problems (configs 1-3): 5 lanes x 10 cells / 10 vehicles, 500 steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dhts_b200 import functional as F, Flags
from dhts_b200.inverse import MacroInverseBatch, MicroInverseBatch, HybridInverseBatch
dev = torch.device("cuda:0")
ev = lambda: torch.cuda.Event(enable_timing=True)
torch.manual_seed(0)
for cls, extra in ((MacroInverseBatch, (10, 5.0)), (MicroInverseBatch, (10, 5.0)), (HybridInverseBatch, (10, 5.0))):
    prob = cls(5, 500, 3, 0.01, 30.0, "x", *extra, device=dev, log_root="/tmp/inv")
    est = prob.initialize()
    x = tuple(s.clone().requires_grad_() for s in est)
    for it in range(3):
        e = [ev() for _ in range(3)]
        for t in x: t.grad = None
        e[0].record(); end = prob.simulate(x, True); e[1].record()
        prob.compute_error(prob.end_state, end).sum().backward(); e[2].record()
        torch.cuda.synchronize()
    t0 = time.time(); prob.num_episode = 10; prob.solve_gd(est); torch.cuda.synchronize(); w = (time.time() - t0) / 10
    print("%-20s fwd %.2f ms  bwd %.2f ms  solve_gd per episode (wall) %.2f ms" % (cls.__name__, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), w * 1e3))
