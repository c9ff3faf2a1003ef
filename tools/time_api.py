"""Wall time of the public numpy API calls on a bench workload (host arrays in, host arrays out)."""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import workloads
from pyresample_b200 import geometry, kd_tree

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = workloads.make_config(cid)
swath = geometry.SwathDefinition(cfg["lon"], cfg["lat"])
area, radius, data = cfg["area"], cfg["radius"], cfg["data"]
warnings.simplefilter("ignore")


def timed(name, fn, n=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("%-52s %8.1f ms" % (name, min(ts)), flush=True)
    return out


info = timed("get_neighbour_info k=1", lambda: kd_tree.get_neighbour_info(swath, area, radius, neighbours=1))
timed("get_sample_from_neighbour_info nn (5 ch)",
      lambda: kd_tree.get_sample_from_neighbour_info('nn', area.shape, data, *info[:3]))
timed("get_neighbour_info k=8", lambda: kd_tree.get_neighbour_info(swath, area, radius * 4, neighbours=8))
timed("resample_nearest (5 ch)", lambda: kd_tree.resample_nearest(swath, data, area, radius))
timed("resample_gauss k=8 (1 ch)", lambda: kd_tree.resample_gauss(swath, data[:, 0], area, radius * 4, radius * 2))
timed("resample_custom k=8 python weight (1 ch)",
      lambda: kd_tree.resample_custom(swath, data[:, 0], area, radius * 4, lambda d: 1 - d / (radius * 4)))
