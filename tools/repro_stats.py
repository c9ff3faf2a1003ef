"""Event counters of the search kernel on a BASELINE configuration (library built with -DPRS_STATS for ONE search
unit, see tools/gpu_stats.sh)."""
import ctypes
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, ".")
import workloads
from pyresample_b200 import _cabi, geometry, kd_tree

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = workloads.make_config(cid)
swath = geometry.SwathDefinition(torch.from_numpy(cfg["lon"]).cuda(), torch.from_numpy(cfg["lat"]).cuda())
data = torch.from_numpy(cfg["data"]).cuda()
warnings.simplefilter("ignore")
L = _cabi.lib()
fn = getattr(ctypes.CDLL(_cabi._SO_PATH), "prs_debug_stats")
out = (ctypes.c_ulonglong * 32)()
fn(out)
if cfg["mode"] == "gauss":
    funcs = [kd_tree._make_gauss(cfg["sigma"])] * cfg["channels"]
    kd_tree._resample(swath, data, cfg["area"], 'custom', cfg["radius"], neighbours=cfg["k"], weight_funcs=funcs)
else:
    src = kd_tree._Source(swath, cfg["area"], True, cfg["radius"], cfg["k"], need_ranks=True)
    kd_tree._query_device(src, cfg["area"], cfg["radius"], cfg["k"], True)
fn(out)
v = [int(x) for x in out]
names = ["valid points", "general: entered", "general: entered full", "general: entered empty", "iter phase0", "iter phase1",
         "iter phase2", "scans phase0", "scans phase1", "scans phase2", "cells phase0", "cells phase1", "cells phase2",
         "general candidates", "tiles staged", "tiles unstaged"] + ["home m=%d" % i for i in range(8)] + \
        ["sent to general", "sent to general full"]
for n, x in zip(names, v):
    print("%-26s %14d  %8.3f per valid point" % (n, x, x / max(v[0], 1)))
