#!/usr/bin/env python
"""bench.py - output Mpixel/s of the kd-tree neighbour resampling path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle), host cores

A "step" is one full pass of the hot path over the workload: validity masks + index build over the source
swath, then the fused query + gather kernel over this rank's rows of the target grid.  N=1 workload:
BASELINE.json configs[1] (AVHRR-like 2048x5400 swath, 5 channels -> global 0.05 deg grid, nearest).
`value` times K steps with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max over
ranks); `e2e` times the public numpy API (pinned host buffers in, host array out).  One JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "output Mpixel/s"


def algorithmic_bytes(n_valid, n_t, channels, coord_bytes=4, data_bytes=4):
    """every valid source point's xyz and data read once + every output written once."""
    return n_valid * (3 * coord_bytes + channels * data_bytes) + n_t * channels * data_bytes


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.active = False  # NVML is initialised before the warm-up (it stalls the driver for milliseconds);
        # samples are only taken while the timed region runs
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop_evt.is_set():
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(float(os.environ.get("PRS_BENCH_CLOCK_PERIOD", "0.05")))
        except Exception as exc:  # NVML missing: report it instead of guessing
            self.reasons.add("nvml_unavailable:%s" % type(exc).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


class CpuReference(object):
    """The reference algorithm on host cores (oracle: numpy ECEF + scipy cKDTree, the reference's own nprocs>1
    backend, _spatial_mp.py:215-226, + numpy gather; restated PROJ inverse included).

    Bounded sample: the full source tree is built ONCE (timed), each step resamples `sample_rows` evenly spaced
    target rows (timed) and scales that to the full grid; a step's figure is N_t / (t_build + t_query_scaled).
    """

    def __init__(self, cfg, sample_rows, workers):
        from oracle import kd_tree_ref, proj_ref
        from oracle.knn_ref import KDTreeRef
        self.cfg, self.workers = cfg, workers
        self.ref, self.proj = kd_tree_ref, proj_ref
        area = cfg["area"]
        self.H, self.W = area.shape
        t0 = time.perf_counter()
        if "source_area" in cfg:
            sa = cfg["source_area"]
            slon, slat = proj_ref.area_lonlats(sa.proj_dict, sa.area_extent, sa.width, sa.height, dtype=sa.dtype)
            slon, slat = slon.ravel(), slat.ravel()
        else:
            slon, slat = cfg["lon"].ravel(), cfg["lat"].ravel()
        with np.errstate(invalid="ignore"):
            valid = (slon >= -180) & (slon <= 180) & (slat <= 90) & (slat >= -90)
        xyz = kd_tree_ref.transform_lonlats(slon[valid], slat[valid])
        self.tree = KDTreeRef(xyz)
        self.tree._get_tree(None)
        self.t_build = time.perf_counter() - t0
        self.dtype = slon.dtype
        self.data = cfg["data"].reshape(slon.size, -1)[valid]
        self.rows = np.unique(np.linspace(0, self.H - 1, sample_rows).astype(int))
        self.n_points = xyz.shape[0]
        self.t_query = None

    def step(self):
        import warnings
        cfg, area, tree = self.cfg, self.cfg["area"], self.tree
        t1 = time.perf_counter()
        n_px = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for r0 in range(0, len(self.rows), 16):
                rr = self.rows[r0:r0 + 16]
                tx, ty = self.proj.proj_vectors(area.area_extent, self.W, self.H, self.dtype)
                gx, gy = np.meshgrid(tx, ty[rr])
                lon, lat = self.proj.inverse(area.proj_dict, gx, gy)
                lon, lat = lon.astype(self.dtype).ravel(), lat.astype(self.dtype).ravel()
                q = self.ref.transform_lonlats(lon, lat)
                dist, idx = tree.query(q, k=cfg["k"], distance_upper_bound=cfg["radius"], workers=self.workers)
                miss = idx == tree.n
                if cfg["mode"] == "nearest":
                    out = self.data[np.where(miss, 0, idx)]
                    out[miss] = 0
                else:
                    w = np.exp(-dist.astype(np.float64) ** 2 / cfg.get("sigma", 25000.0) ** 2)
                    w[miss] = 0
                    vals = self.data[np.where(miss, 0, idx)]
                    out = (w[..., None] * vals).sum(1) / np.maximum(w.sum(1), 1e-300)[:, None]
                n_px += lon.size
        self.t_query = time.perf_counter() - t1
        t_full = self.t_build + self.t_query * (self.H * self.W / n_px)
        return self.H * self.W / t_full / 1e6

    def describe(self):
        return ("source tree over all %d points built once (%.1f s, included in every step) + %d of %d target rows "
                "per step (%.2f s) scaled to the full grid; numpy ECEF + scipy cKDTree(leafsize=16, workers=%d) + "
                "numpy gather; restated PROJ inverse included"
                % (self.n_points, self.t_build, len(self.rows), self.H, self.t_query or 0.0, self.workers))


def run_reference(args, rank, world):
    if rank != 0:
        return
    import workloads
    cfg = workloads.make_config(args.config, scale=args.scale)
    workers = os.cpu_count() or 1
    ref = CpuReference(cfg, args.ref_rows, workers)
    for _ in range(args.warmup):
        ref.step()
    vals = [ref.step() for _ in range(args.steps)]
    value = float(np.mean(vals))
    H, W = cfg["area"].shape
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": H * W / value / 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["name"], "scale": args.scale},
            "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": workers, "kind": "port",
                             "sample": ref.describe()},
            "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configs index + 1 (default 2 = configs[1])")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--ref-rows", type=int, default=96, help="target rows sampled by the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import workloads
    from pyresample_b200 import _cabi, distributed, geometry, kd_tree
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout; stdout carries ONE JSON line
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _cabi.lib()
    # every rank derives the (seeded) geometry; the source ARRAYS that are used come from rank 0 via NCCL
    cfg = workloads.make_config(args.config, scale=args.scale)
    area = cfg["area"]
    H, W = area.shape
    channels = cfg["channels"]
    rows = distributed.cyclic_rows(H, world, rank)  # stripes of 64 rows dealt round-robin
    nrows = rows[1]

    # ---- inputs: pinned host copies (e2e) and HBM-resident tensors (value); source broadcast once from rank 0
    host = {}
    dev = {}
    if "source_area" in cfg:
        names = ("data",)
    else:
        names = ("lon", "lat", "data")
    for name in names:
        if rank == 0:
            arr = np.ascontiguousarray(cfg[name])
            pin = torch.empty(arr.shape, dtype=getattr(torch, str(arr.dtype)), pin_memory=True)
            pin.numpy()[...] = arr
            host[name] = pin
            dev[name] = pin.cuda(non_blocking=True)
        else:
            dev[name] = torch.empty(cfg[name].shape, dtype=getattr(torch, str(cfg[name].dtype)), device="cuda")
    torch.cuda.synchronize()
    if world > 1:
        distributed.broadcast_source([dev[n] for n in names])
        torch.cuda.synchronize()
    if "source_area" in cfg:
        source_dev = cfg["source_area"]
        n_src = source_dev.size
    else:
        source_dev = geometry.SwathDefinition(dev["lon"], dev["lat"])
        n_src = dev["lon"].numel()
    data_dev = dev["data"]
    radius, k, mode = cfg["radius"], cfg["k"], cfg["mode"]
    sigmas = [cfg.get("sigma", 25000.0)] * channels if channels > 1 else cfg.get("sigma", 25000.0)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    q_events = []
    b_events = []
    info = {}

    fill = torch.zeros((channels,), dtype=data_dev.dtype, device="cuda")  # fill_value = 0: a constant of the call

    def step(record=False):
        """One pass of the hot path on resident inputs: masks + index build, then fused query + gather."""
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        source = kd_tree._Source(source_dev, area, True, radius, k, need_ranks=False,
                                 cover_rows=rows if world > 1 else None)
        e1.record()
        tg, n_t, keep = kd_tree._make_target(source, area, True, radius, rows=rows)
        stream = kd_tree._stream()
        if mode == "nearest":
            out = torch.empty((n_t, channels), dtype=data_dev.dtype, device="cuda")
            _cabi.check(L.prs_resample_nearest(source.index.handle, ctypes.byref(tg), float(radius),
                                               kd_tree._ptr(data_dev), channels * data_dev.element_size(),
                                               kd_tree._ptr(fill), kd_tree._ptr(out), stream))
        else:
            acc = torch.float64 if channels > 1 else torch.float32
            out = torch.empty((n_t, channels), dtype=acc, device="cuda")
            sig = (ctypes.c_double * channels)(*([cfg.get("sigma", 25000.0)] * channels))
            _cabi.check(L.prs_resample_gauss(source.index.handle, ctypes.byref(tg), int(k), float(radius),
                                             kd_tree._ptr(data_dev), 0, channels, sig, 0.0, kd_tree._ptr(out), None,
                                             None, None, stream))
        e2.record()
        if record:
            b_events.append((e0, e1))
            q_events.append((e1, e2))
        info["n_valid"] = int(source.n_valid)
        info["n_indexed"] = int(source.index.n_indexed)
        info["cells_per_side"] = int(source.index.cells_per_side)
        info["index_bytes"] = int(source.index.bytes)
        return out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    out = None
    for _ in range(args.warmup):
        out = step()  # held like in the timed loop, so that torch's allocator already caches both output buffers
    sync_all()
    sampler.active = True
    launches0 = L.prs_launch_count()
    t_start, t_end = ev(), ev()
    t_start.record()
    for _ in range(args.steps):
        out = step(record=True)
    t_end.record()
    sync_all()
    launches = L.prs_launch_count() - launches0
    if not sampler.samples:  # a very short timed region: take one sample right at its end, still under load
        time.sleep(float(os.environ.get("PRS_BENCH_CLOCK_PERIOD", "0.05")) * 1.5)
    clocks = sampler.stop()
    elapsed_ms = t_start.elapsed_time(t_end)
    q_all = [a.elapsed_time(b) for a, b in q_events]
    if os.environ.get("PRS_BENCH_DEBUG"):
        print("query ms per step:", " ".join("%.3f" % v for v in q_all), file=sys.stderr)
        print("build ms per step:", " ".join("%.3f" % a.elapsed_time(b) for a, b in b_events), file=sys.stderr)
    b_all = [a.elapsed_time(b) for a, b in b_events]
    q_ms, b_ms = float(np.mean(q_all)), float(np.mean(b_all))
    if world > 1:
        t = torch.tensor([elapsed_ms, q_ms, b_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, q_ms, b_ms = [float(v) for v in t.tolist()]
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = elapsed_ms / args.steps
    n_t_total = H * W
    value = n_t_total / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API: pinned host inputs -> H2D -> hot path -> D2H of the result
    e2e = None
    if not args.no_e2e:
        h2d = sum(host[n].numel() * host[n].element_size() for n in names) if rank == 0 else 0
        itemsize = 8 if (mode != "nearest" and channels > 1) else data_dev.element_size()
        d2h = n_t_total * channels * itemsize

        def e2e_step():
            if world == 1:
                if "source_area" in cfg:
                    src = cfg["source_area"]
                else:
                    src = geometry.SwathDefinition(host["lon"].numpy(), host["lat"].numpy())
                if mode == "nearest":
                    return kd_tree.resample_nearest(src, host["data"].numpy(), area, radius)
                return kd_tree.resample_gauss(src, host["data"].numpy(), area, radius, sigmas, neighbours=k)
            # N > 1: rank 0 uploads, NCCL broadcast, every rank resamples and reads back its own rows
            if mode == "nearest" and "source_area" not in cfg:
                # pipelined: coordinates first, the data follow while every rank builds its index and fills its rows
                return distributed.resample_nearest_from_host(
                    host["lon"].numpy() if rank == 0 else None, host["lat"].numpy() if rank == 0 else None,
                    host["data"].numpy() if rank == 0 else None, area, radius, n_src, cfg["lon"].dtype,
                    cfg["data"].dtype, channels)
            for n in names:
                if rank == 0:
                    dev[n].copy_(host[n], non_blocking=True)
            distributed.broadcast_source([dev[n] for n in names])
            src = cfg["source_area"] if "source_area" in cfg else geometry.SwathDefinition(dev["lon"], dev["lat"])
            if mode == "nearest":
                blk = distributed.resample_nearest_sharded(src, dev["data"], area, radius)
            else:
                blk = kd_tree._resample(src, dev["data"], area, 'custom', radius, neighbours=k,
                                        weight_funcs=[kd_tree._make_gauss(s) for s in sigmas] if channels > 1
                                        else kd_tree._make_gauss(sigmas), _rows=rows)
            return kd_tree._host(blk)

        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(max(4, min(args.warmup, 6))):  # first calls grow the pinned staging buffers
                e2e_step()
            sync_all()
            n_e2e = max(2, min(args.steps, 5))
            per_call = []
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                t1 = time.perf_counter()
                e2e_step()
                per_call.append((time.perf_counter() - t1) * 1e3)
            sync_all()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
            if os.environ.get("PRS_BENCH_DEBUG"):
                print("e2e ms per call:", " ".join("%.2f" % v for v in per_call), file=sys.stderr)
        if world > 1:
            t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": n_t_total / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "api": ("pyresample_b200.kd_tree.resample_%s (numpy in, numpy out)" % mode) if world == 1 else
                      ("pyresample_b200.distributed.resample_nearest_from_host (host arrays on rank 0 in, each rank's "
                       "rows out to its host)" if mode == "nearest" and "source_area" not in cfg else
                       "rank 0 uploads, NCCL broadcast, kd_tree resample of each rank's rows, read-back")}

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        n_t_rank = nrows * W
        itemsize = data_dev.element_size()
        # rank 0's share: the source points it indexes (all valid points at N = 1) and its target rows
        b_alg = algorithmic_bytes(info["n_indexed"], n_t_rank, channels, 4, itemsize)
        achieved = b_alg / (q_ms * 1e-3) / 1e9
        traffic = None  # dram__bytes_read + write of the same kernels from the committed ncu --set full capture
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                tj = json.load(fh).get("config%d" % args.config)
            if tj and world == 1 and args.scale == 1.0:
                traffic = tj["dram_bytes_per_step"]
        except Exception:
            pass
        roofline = {"bound": "hbm",
                    "kernel": "prs::classify_kernel + prs::query_kernel (fused target transform + k-NN + gather)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_kind": peak_kind, "kernel_ms": q_ms, "algorithmic_bytes_per_launch": int(b_alg),
                    "index_build_ms": b_ms, "kernel_ms_min": float(np.min(q_all)),
                    "kernel_ms_median": float(np.median(q_all)), "kernel_mpixel_per_s": n_t_rank / (q_ms * 1e-3) / 1e6}
        line = {"metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["name"], "scale": args.scale, "source_points": int(n_src),
                           "valid_source_points": info["n_valid"], "target_pixels": int(n_t_total),
                           "channels": channels, "neighbours": k, "radius_m": radius,
                           "sharding": ("target rows in stripes of %d dealt round-robin over %d rank(s); source broadcast once "
                                        "(NCCL), each rank indexes only the source points near its rows"
                                        % (distributed.DEFAULT_STRIPE, world)),
                           "l2": "inputs (%.0f MB index + data, %.0f MB output) exceed the 126 MB L2; no flush needed"
                                 % ((info["index_bytes"] + data_dev.numel() * itemsize) / 1e6,
                                    n_t_total * channels * itemsize / 1e6),
                           "cells_per_face_side": info["cells_per_side"]},
                "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e}
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(cfg, args.ref_rows, os.cpu_count() or 1)
            v = float(np.mean([ref.step() for _ in range(3)]))
            line["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": ref.describe()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
