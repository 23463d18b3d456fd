"""Phase-by-phase cycle breakdown of the hybrid network kernels on one config-4 episode (R = 1).
Needs a diagnosis build:  DHTS_NVCC_EXTRA=-DDHTS_PHASE_TIMING python -c 'from dhts_b200 import _build; _build.build(force=True)'
usage: python scripts/hyb_phases.py   (run on a GPU box; rebuild without the flag afterwards)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from dhts_b200 import _lib  # noqa: E402
from itscp_env_cases import c4_env, c4_fixture, c4_spawn_routes  # noqa: E402

lib = ctypes.CDLL(_lib.SO_PATH)
buf = (ctypes.c_ulonglong * 32)()
dev = torch.device("cuda", 0)
G = c4_fixture()
env = c4_env(G, dev)
sp = torch.tensor(c4_spawn_routes(G, env.topo), dtype=torch.int32, device=dev)
NAMES = {0: "fwd S+0 heads, aux copy", 1: "fwd ghosts + micro lanes", 2: "fwd fluxes", 3: "fwd update + walk lookups",
         4: "fwd candidates", 5: "fwd conversions", 10: "bwd rows landed", 11: "bwd replay tail", 12: "bwd R1 conversions",
         13: "bwd R2a fold/ghosts/micro", 14: "bwd R2b interfaces", 15: "bwd R3 gathers", 16: "bwd R2c cells/sides"}
for it in range(3):
    action = torch.tensor(G["action"], dtype=torch.float64, device=dev, requires_grad=True)
    reward = env.rollout(action[None], True, spawn_routes=sp)[0]
    torch.cuda.synchronize()
    lib.dhts_debug_phase_cycles(buf); fwd = list(buf)
    reward.backward()
    torch.cuda.synchronize()
    lib.dhts_debug_phase_cycles(buf); bwd = list(buf)
T = int(G["T"])
print("forward kernel, cycles per step (thread 0 of CTA 0):")
for i in range(6):
    print("  %-32s %8.0f" % (NAMES[i], fwd[i] / T))
print("  (cost of one marker: %.0f cycles; steps with a flagged conversion group: %d of %d)" % (fwd[6] / T, fwd[20], T))
print("  total %.0f cycles / step = %.2f us at 1.965 GHz" % (sum(fwd[:6]) / T, sum(fwd[:6]) / T / 1965.0))
print("adjoint kernel, cycles per step (replay phases 0-5 listed first):")
for i in list(range(6)) + [10, 11, 12, 13, 14, 16, 15]:
    print("  %-32s %8.0f" % (NAMES[i], bwd[i] / T))
print("  total %.0f cycles / step = %.2f us" % (sum(bwd) / T, sum(bwd) / T / 1965.0))
