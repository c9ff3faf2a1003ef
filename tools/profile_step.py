"""Host-side cProfile of the resident-input step (tensor API) on the bench workload."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import workloads
from pyresample_b200 import geometry, kd_tree

cfg = workloads.make_config(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
lon, lat, data = (torch.from_numpy(cfg[k]).cuda() for k in ("lon", "lat", "data"))
src = geometry.SwathDefinition(lon, lat)


def call():
    return kd_tree.resample_nearest(src, data, cfg["area"], cfg["radius"])


for _ in range(5):
    call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    call()
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    call()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
