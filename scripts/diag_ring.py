import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dhts_b200
from dhts_b200 import functional as F
dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
for dtype in (torch.float64, torch.float32):
    for B, N, T in ((701, 1024, 7), (301, 64, 9), (9, 256, 2), (40, 2048, 5), (3, 1024, 24)):
        r0 = rng.uniform(0, 1, (B, N)); u0 = rng.uniform(0, 30, (B, N))
        gr = rng.uniform(0, 1, (B, 2)); gu = rng.uniform(0, 30, (B, 2)); w = rng.normal(size=(B, N))
        t = lambda a: torch.tensor(a, dtype=dtype, device=dev)
        out = {}
        for ring in ("4", "2", "0", "0b"):
            os.environ["DHTS_ARZ_RING"] = ring[0]
            flags = dhts_b200.Flags(dev)
            tr, tu = t(r0).requires_grad_(), t(u0).requires_grad_()
            tgr, tgu = t(gr).requires_grad_(), t(gu).requires_grad_()
            rT, yT, uT = F.arz_rollout(tr, tu, tgr, tgu, 5.0, 30.0, 0.01, T, ckpt_every=1, flags=flags)
            ((rT * t(w)).sum() + (uT * t(w / 30)).sum()).backward()
            print("flags", flags.t.tolist(), end=" ")
            out[ring] = [x.clone() for x in (tr.grad, tu.grad, tgr.grad, tgu.grad)]
        for ring in ("4", "2", "0b"):
            d = [float((a - b).abs().max()) for a, b in zip(out[ring], out["0"])]
            m = [float(b.abs().max()) for b in out["0"]]
            bad = (out[ring][0] != out["0"][0]).nonzero()
            print("\n", dtype, (B, N, T), "ring", ring, "maxabs diff", d, "scale", m, "n_bad", len(bad), bad[:6].tolist())
