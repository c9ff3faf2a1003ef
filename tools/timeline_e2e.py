"""Host-side time stamps of the pipelined resample_nearest on the bench workload (ms since entering the call)."""
import os
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
for _ in range(4):
    kd_tree.resample_nearest(src, host["data"], cfg["area"], cfg["radius"])
kd_tree._TIMELINE = []
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kd_tree.resample_nearest(src, host["data"], cfg["area"], cfg["radius"])
    print("call %.2f ms" % ((time.perf_counter() - t0) * 1e3), kd_tree._TIMELINE[-1])
# raw PCIe rates for reference
d = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
h = torch.empty(512 << 20, dtype=torch.uint8, pin_memory=True)
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (a.numel() / (time.perf_counter() - t0) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)
h2 = torch.empty(512 << 20, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s1):
    d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2):
    h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("duplex: 2 x 512 MiB in %.2f ms -> %.1f GB/s each way" % ((time.perf_counter() - t0) * 1e3,
                                                                 d.numel() / (time.perf_counter() - t0) / 1e9))
