"""Per-step wall/device times of the resident step, to spot intermittent slow steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import workloads
from pyresample_b200 import geometry, kd_tree

cfg = workloads.make_config(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
lon, lat, data = (torch.from_numpy(cfg[k]).cuda() for k in ("lon", "lat", "data"))
src = geometry.SwathDefinition(lon, lat)
times = []
for i in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = kd_tree.resample_nearest(src, data, cfg["area"], cfg["radius"])
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
print(" ".join("%.2f" % t for t in times))
print("mem", torch.cuda.mem_get_info())
