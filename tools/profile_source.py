"""Host-side cProfile of the index build (_Source) on a bench workload: finds Python time between kernel launches."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import workloads
from pyresample_b200 import geometry, kd_tree

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = workloads.make_config(cid)
if "source_area" in cfg:
    src = cfg["source_area"]
else:
    lon, lat = (torch.from_numpy(cfg[k]).cuda() for k in ("lon", "lat"))
    src = geometry.SwathDefinition(lon, lat)
k = cfg.get("k", 1)


def call():
    return kd_tree._Source(src, cfg["area"], True, cfg["radius"], k, need_ranks=False)


for _ in range(5):
    call()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = call()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("ms per build:", " ".join("%.2f" % t for t in ts))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    call()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
