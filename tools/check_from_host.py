"""Run under torchrun (N >= 1 GPUs): distributed.resample_nearest_from_host on every rank must equal the rows
of the single-GPU resample_nearest that belong to the rank.  Prints one OK line per rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import workloads
from pyresample_b200 import distributed, geometry, kd_tree

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
cfg = workloads.make_config(2, scale=float(sys.argv[1]) if len(sys.argv) > 1 else 0.25)
area = cfg["area"]
lon, lat, data = cfg["lon"], cfg["lat"], cfg["data"]
ch = data.shape[1]
mine = distributed.resample_nearest_from_host(lon if rank == 0 else None, lat if rank == 0 else None,
                                              data if rank == 0 else None, area, cfg["radius"], lon.size, lon.dtype,
                                              data.dtype, ch, fill_value=-5)
full = kd_tree.resample_nearest(geometry.SwathDefinition(lon, lat), data, area, cfg["radius"], fill_value=-5)
rows = distributed.cyclic_global_rows(area.shape[0], world, rank, distributed.DEFAULT_STRIPE)
want = full[rows]
ok = mine.shape == want.shape and np.array_equal(mine, want, equal_nan=True)
print("rank %d/%d rows %d %s, found %.1f%%" % (rank, world, len(rows), "OK" if ok else "MISMATCH",
                                               100.0 * np.mean(want[..., 0] != -5)), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
