"""cProfile of one e2e resample_nearest call on the bench workload (host-side overhead hunting)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import workloads
from pyresample_b200 import geometry, kd_tree

cfg = workloads.make_config(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
host = {}
for name in ("lon", "lat", "data"):
    arr = np.ascontiguousarray(cfg[name])
    pin = torch.empty(arr.shape, dtype=getattr(torch, str(arr.dtype)), pin_memory=True)
    pin.numpy()[...] = arr
    host[name] = pin.numpy()
src = geometry.SwathDefinition(host["lon"], host["lat"])


def call():
    return kd_tree.resample_nearest(src, host["data"], cfg["area"], cfg["radius"])


for _ in range(3):
    call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    call()
torch.cuda.synchronize()
print("ms per call", (time.perf_counter() - t0) / 3 * 1e3)
pr = cProfile.Profile()
pr.enable()
call()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
