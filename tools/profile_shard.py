"""One rank's share of the sharded step on a single GPU (what rank R of W ranks runs in bench.py --gpus W),
for per-kernel launch lists: python tools/profile_shard.py [config] [world] [rank]."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import workloads
from pyresample_b200 import _cabi, distributed, geometry, kd_tree

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = workloads.make_config(cid)
area = cfg["area"]
lon, lat, data = (torch.from_numpy(cfg[k]).cuda() for k in ("lon", "lat", "data"))
src = geometry.SwathDefinition(lon, lat)
rows = distributed.cyclic_rows(area.shape[0], world, rank)
L = _cabi.lib()
ch = data.shape[1]
fill = torch.zeros((ch,), dtype=data.dtype, device="cuda")


def step():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    source = kd_tree._Source(src, area, True, cfg["radius"], 1, need_ranks=False, cover_rows=rows if world > 1 else None)
    e[1].record()
    tg, n_t, keep = kd_tree._make_target(source, area, True, cfg["radius"], rows=rows)
    out = torch.empty((n_t, ch), dtype=data.dtype, device="cuda")
    _cabi.check(L.prs_resample_nearest(source.index.handle, ctypes.byref(tg), float(cfg["radius"]),
                                       kd_tree._ptr(data), ch * data.element_size(), kd_tree._ptr(fill),
                                       kd_tree._ptr(out), kd_tree._stream()))
    e[2].record()
    return e, out, int(source.index.n_indexed)


for _ in range(5):
    step()
torch.cuda.synchronize()
res = []
for _ in range(20):
    e, out, n_idx = step()
    torch.cuda.synchronize()
    res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
b, q = np.median([r[0] for r in res]), np.median([r[1] for r in res])
print("config %d rank %d/%d: rows %d, indexed %d points, build %.3f ms, query %.3f ms" %
      (cid, rank, world, out.shape[0] // area.shape[1], n_idx, b, q))
