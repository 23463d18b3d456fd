"""Times the connected-network legs of bench.py (itscp_net at several replica counts; itscp_c4) without the headline pass.
usage: python scripts/net_time.py [replicas ...]"""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
for R in ([int(x) for x in sys.argv[1:]] or [2048, 256, 1]):
    a = types.SimpleNamespace(net_replicas=R)
    d = bench.network_bench(a, dev, torch.float64, torch)
    out["itscp_net_R%d" % R] = {k: d[k] for k in ("value", "fwd_ms", "bwd_ms", "mean_reward", "grad_abs_mean")}
try:
    d = bench.config4_bench(types.SimpleNamespace(net_replicas=2048), dev, torch.float64, torch)
    out["itscp_c4"] = {k: v for k, v in d.items() if k != "workload"}
except Exception as e:      # noqa: BLE001
    out["itscp_c4"] = repr(e)
print(json.dumps(out, indent=1))
