import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import workloads
from pyresample_b200 import kd_tree, distributed, geometry
cfg = workloads.make_config(5)
src, tgt, radius = cfg["source_area"], cfg["area"], cfg["radius"]
H = tgt.shape[0]
lon_full, lat_full, _ = kd_tree._geo_lonlats_device(src)
valid = (~torch.isnan(lon_full)) & (lon_full.abs() <= 180)
print("valid source px", int(valid.sum()), "of", lon_full.numel())
for world in (8, 4):
    rows = distributed.cyclic_rows(H, world, 0)
    cover = kd_tree._target_cover(tgt, rows, radius, np.dtype(src.dtype))
    print("world", world, "rows", rows, "cover cells set", int(cover.sum()), "of", cover.numel())
    lon, lat, _ = kd_tree._area_lonlats_covered(src, cover)
    there = ~torch.isnan(lon)
    print("  transformed px", int((there & valid).sum()), "dropped valid px", int((~there & valid).sum()))
    for name, fn in (("plain", lambda: kd_tree._geo_lonlats_device(src)), ("covered", lambda: kd_tree._area_lonlats_covered(src, cover))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print("  %s %.3f ms" % (name, e0.elapsed_time(e1) / 10))
    s0 = kd_tree._Source(src, tgt, True, radius, 1, need_ranks=False, cover_rows=rows)
    print("  n_indexed", s0.index.n_indexed)
