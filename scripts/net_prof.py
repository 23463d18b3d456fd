"""Small driver for profiling the connected-network kernels: the bench's secondary sections alone."""
import argparse
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--net-replicas", type=int, default=592)
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
dev = torch.device("cuda:0")
dt_t = torch.float64 if a.dtype == "f64" else torch.float32
print(bench.network_bench(a, dev, dt_t, torch))
print(bench.config4_bench(a, dev, dt_t, torch))
