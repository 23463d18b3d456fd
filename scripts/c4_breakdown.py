"""Where one config-4 episode (R = 1 and R = 256) spends its time: fused forward, reward graph, adjoint."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dhts_b200.itscp_env import ItscpEnv, problem_1
import dhts_b200.hybrid_network as H

dev = torch.device("cuda:0")
np.random.seed(1)
env = ItscpEnv(device=dev); env.schedule_callback = problem_1
env.config.update(num_intersection=3, lane_length=5.0, num_lane=1, policy_length=20, signal_length=4, mode="hybrid", speed_limit=60.0)
env.reset()
ev = lambda: torch.cuda.Event(enable_timing=True)
orig_fwd, orig_bwd = H.HybRolloutFn.forward, H.HybRolloutFn.backward
T = {}
def timed(name, fn):
    def w(*a, **k):
        e0, e1 = ev(), ev(); e0.record(); out = fn(*a, **k); e1.record(); T.setdefault(name, []).append((e0, e1)); return out
    return w
H.HybRolloutFn.forward = staticmethod(timed("kernel_fwd", orig_fwd))
H.HybRolloutFn.backward = staticmethod(timed("kernel_bwd", orig_bwd))
g = torch.Generator().manual_seed(3)
for R in (1, 256, 2048):
    act = (0.3 + 0.4 * torch.rand((R, 45), generator=g)).double().to(dev).requires_grad_()
    env.resample_spawn_routes(R, g)
    for it in range(4):
        T.clear(); act.grad = None
        e = [ev() for _ in range(3)]
        e[0].record(); r = env.rollout(act, True); e[1].record(); r.sum().backward(); e[2].record()
        torch.cuda.synchronize()
    kf = sum(a.elapsed_time(b) for a, b in T["kernel_fwd"]); kb = sum(a.elapsed_time(b) for a, b in T["kernel_bwd"])
    print("R=%d rollout(total fwd) %.2f ms [hyb fwd fn %.2f]  backward(total) %.2f ms [hyb bwd fn %.2f]" %
          (R, e[0].elapsed_time(e[1]), kf, e[1].elapsed_time(e[2]), kb))
