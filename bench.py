#!/usr/bin/env python
"""bench.py - output Mpixel/s of the kd-tree neighbour resampling path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle), host cores

A "step" is one full pass of the hot path over the workload: validity masks + index build over the source,
then the query + gather / weighting over this rank's rows of the target grid, exactly as the configuration says:

    nearest (C1, C2, C5)  prs_index_build + prs_resample_nearest (fused search + gather)
    gauss   (C3)          prs_index_build + prs_resample_gauss   (fused search + Gaussian weighting, f64 accumulation)
    custom  (C4)          kd_tree.resample_custom's device path: prs_index_build + prs_query (k = 16 neighbour info)
                          + the Python weight function `1 - d / 1e5` evaluated on a CUDA view of the distances
                          + prs_gather_weighted; XArrayResamplerNN is timed next to it (`xarray_nn` sub-record)

N=1 workload: BASELINE.json configs[1] (C2: AVHRR-like 2048x5400 swath, 5 channels -> global 0.05 deg grid, nearest).
`value` times K steps with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max over
ranks); `e2e` times the public numpy API (pinned host buffers in, host array out).  `configs` carries the same
resident measurement for C5, the configuration BASELINE.json names for the 8-GPU run, at the run's N.
One JSON line on rank 0.
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


def custom_weight(d):
    """The reference tests' weight function (test_kd_tree.py:74-75), the one SURVEY 8(d) names for C4."""
    return 1 - d / 1e5


def algorithmic_bytes(n_valid, n_t, channels, coord_bytes=4, data_bytes=4, out_bytes=None):
    """every valid source point's xyz and data read once + every output written once."""
    out_bytes = data_bytes if out_bytes is None else out_bytes
    return n_valid * (3 * coord_bytes + channels * data_bytes) + n_t * channels * out_bytes


def workload_config(cfg, scale):
    """The part of `config` both arms share (the driver compares it between the arms)."""
    area = cfg["area"]
    n_src = cfg["source_area"].size if "source_area" in cfg else cfg["lon"].size
    return {"workload": cfg["name"], "scale": scale, "source_points": int(n_src),
            "target_pixels": int(area.shape[0] * area.shape[1]), "channels": int(cfg["channels"]),
            "neighbours": int(cfg["k"]), "radius_m": float(cfg["radius"]), "mode": cfg["mode"]}


def arithmetic_dtype(cfg):
    """Type the path computes in: float32 coordinates / distances; multi-channel weighting accumulates in float64
    (numpy promotion in _resample_with_weights, kd_tree.py:783)."""
    if cfg["mode"] == "nearest":
        return "f32"
    return "f64" if cfg["channels"] > 1 else "f32"


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
    backend, _spatial_mp.py:215-226, + numpy gather / weighting; restated PROJ inverse included).

    A step is what one reference call does: range mask + ECEF + tree build over the source, then every `rows`-th
    part of the target resampled in blocks of 16 rows.  With `rows = all` nothing is extrapolated; otherwise the
    query part is scaled to the full grid and the figure is N_t / (t_build + t_query_scaled)."""

    def __init__(self, cfg, workers):
        from oracle import kd_tree_ref, proj_ref
        from oracle.knn_ref import KDTreeRef
        self.cfg, self.workers = cfg, workers
        self.ref, self.proj, self.KDTreeRef = kd_tree_ref, proj_ref, KDTreeRef
        self.H, self.W = cfg["area"].shape
        self.rows = np.arange(self.H)
        self.t_build = self.t_query = None
        self.n_points = 0

    def set_rows(self, n_rows):
        n_rows = int(max(16, min(self.H, n_rows)))
        self.rows = np.arange(self.H) if n_rows >= self.H else \
            np.unique(np.linspace(0, self.H - 1, n_rows).astype(int))

    def build(self):
        cfg = self.cfg
        t0 = time.perf_counter()
        if "source_area" in cfg:
            sa = cfg["source_area"]
            slon, slat = self.proj.area_lonlats(sa.proj_dict, sa.area_extent, sa.width, sa.height, dtype=sa.dtype)
            slon, slat = slon.ravel(), slat.ravel()
        else:
            slon, slat = cfg["lon"].ravel(), cfg["lat"].ravel()
        with np.errstate(invalid="ignore"):
            valid = (slon >= -180) & (slon <= 180) & (slat <= 90) & (slat >= -90)
        xyz = self.ref.transform_lonlats(slon[valid], slat[valid])
        tree = self.KDTreeRef(xyz)
        tree._get_tree(None)
        data = cfg["data"].reshape(slon.size, -1)[valid]
        self.t_build = time.perf_counter() - t0
        self.n_points = xyz.shape[0]
        return tree, data, slon.dtype

    def step(self):
        import warnings
        cfg, area = self.cfg, self.cfg["area"]
        tree, data, dtype = self.build()
        t1 = time.perf_counter()
        n_px = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tx, ty = self.proj.proj_vectors(area.area_extent, self.W, self.H, dtype)
            for r0 in range(0, len(self.rows), 16):
                rr = self.rows[r0:r0 + 16]
                gx, gy = np.meshgrid(tx, ty[rr])
                lon, lat = self.proj.inverse(area.proj_dict, gx, gy)
                lon, lat = lon.astype(dtype).ravel(), lat.astype(dtype).ravel()
                q = self.ref.transform_lonlats(lon, lat)
                dist, idx = tree.query(q, k=cfg["k"], distance_upper_bound=cfg["radius"], workers=self.workers)
                miss = idx == tree.n
                if cfg["mode"] == "nearest":
                    out = data[np.where(miss, 0, idx)]
                    out[miss] = 0
                else:
                    if cfg["mode"] == "gauss":
                        w = np.exp(-dist.astype(np.float64) ** 2 / cfg.get("sigma", 25000.0) ** 2)
                    else:
                        w = custom_weight(dist.astype(np.float64))
                    w[miss] = 0
                    vals = data[np.where(miss, 0, idx)]
                    out = (w[..., None] * vals).sum(1) / np.maximum(w.sum(1), 1e-300)[:, None]
                n_px += lon.size
        self.t_query = time.perf_counter() - t1
        t_full = self.t_build + self.t_query * (self.H * self.W / n_px)
        self.t_step = self.t_build + self.t_query
        return self.H * self.W / t_full / 1e6

    def describe(self):
        what = "every target row" if len(self.rows) >= self.H else \
            "%d of %d target rows, query time scaled to the full grid" % (len(self.rows), self.H)
        return ("each step: range mask + ECEF + cKDTree(leafsize=16) build over all %d source points (%.1f s, one "
                "thread like pykdtree's build) + %s (%.2f s): restated PROJ inverse, numpy ECEF, "
                "cKDTree.query(workers=%d), numpy gather / weighting"
                % (self.n_points, self.t_build or 0.0, what, self.t_query or 0.0, self.workers))


def run_reference(args, rank, world):
    if rank != 0:
        return
    import workloads
    cfg = workloads.make_config(args.config, scale=args.scale)
    workers = os.cpu_count() or 1
    ref = CpuReference(cfg, workers)
    H, W = cfg["area"].shape
    # size the per-step sample so that the whole run fits the time budget: a probe of 48 rows gives the cost per row
    budget = float(os.environ.get("PRS_REF_BUDGET_S", "300"))
    if args.ref_rows > 0:
        ref.set_rows(args.ref_rows)
    else:
        ref.set_rows(48)
        ref.step()
        per_row = ref.t_query / len(ref.rows)
        per_step = budget / max(1, args.steps + args.warmup) - ref.t_build
        ref.set_rows(H if per_step >= per_row * H else max(16, per_step / per_row))
    for _ in range(args.warmup):
        ref.step()
    vals, wall = [], []
    for _ in range(args.steps):
        vals.append(ref.step())
        wall.append(ref.t_step)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(wall)) * 1e3,
            "ms_per_full_workload": H * W / value / 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": arithmetic_dtype(cfg),
            "data": "synthetic", "config": workload_config(cfg, args.scale),
            "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": workers, "kind": "port",
                             "sample": ref.describe()},
            "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class Resident(object):
    """One configuration with its inputs resident in HBM: `step()` is one pass of the hot path on this rank's rows."""

    def __init__(self, cid, scale, rank, world):
        import torch

        import workloads
        from pyresample_b200 import distributed, geometry
        self.torch, self.cid = torch, cid
        self.rank, self.world = rank, world
        self.cfg = cfg = workloads.make_config(cid, scale=scale)
        self.area = cfg["area"]
        self.H, self.W = self.area.shape
        self.channels, self.k, self.mode, self.radius = cfg["channels"], cfg["k"], cfg["mode"], cfg["radius"]
        self.rows = distributed.cyclic_rows(self.H, world, rank)  # stripes of 64 rows dealt round-robin
        self.names = ("data",) if "source_area" in cfg else ("lon", "lat", "data")
        # pinned host copies on rank 0 (the e2e inputs) and HBM-resident tensors; the source arrays every rank
        # uses come from rank 0 through NCCL
        self.host, self.dev = {}, {}
        for name in self.names:
            if rank == 0:
                arr = np.ascontiguousarray(cfg[name])
                pin = torch.empty(arr.shape, dtype=getattr(torch, str(arr.dtype)), pin_memory=True)
                pin.numpy()[...] = arr
                self.host[name] = pin
                self.dev[name] = pin.cuda(non_blocking=True)
            else:
                self.dev[name] = torch.empty(cfg[name].shape, dtype=getattr(torch, str(cfg[name].dtype)), device="cuda")
        torch.cuda.synchronize()
        if world > 1:
            distributed.broadcast_source([self.dev[n] for n in self.names])
            torch.cuda.synchronize()
        if "source_area" in cfg:
            self.source_dev = cfg["source_area"]
            self.n_src = self.source_dev.size
        else:
            self.source_dev = geometry.SwathDefinition(self.dev["lon"], self.dev["lat"])
            self.n_src = self.dev["lon"].numel()
        self.data_dev = self.dev["data"]
        self.fill = torch.zeros((self.channels,), dtype=self.data_dev.dtype, device="cuda")
        self.info = {}
        self.q_events, self.b_events = [], []
        self.weight_funcs = [custom_weight] * self.channels if self.channels > 1 else custom_weight

    def step(self, record=False, want_info=False):
        from pyresample_b200 import _cabi, kd_tree
        torch, L = self.torch, _cabi.lib()
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        source = kd_tree._Source(self.source_dev, self.area, True, self.radius, self.k,
                                 need_ranks=self.mode == "custom",
                                 cover_rows=self.rows if self.world > 1 else None,
                                 crop_to_cover=self.mode == "nearest" and self.world > 1)
        e1.record()
        channels, data_dev = self.channels, self.data_dev
        if self.mode == "custom":
            out = kd_tree._custom_on_device(source, self.area, self.radius, self.k, True,
                                            data_dev.reshape(self.n_src, -1), np.dtype(np.float32), channels,
                                            self.weight_funcs, channels > 1, 0.0, False, rows=self.rows)[0]
        else:
            tg, n_t, keep = kd_tree._make_target(source, self.area, True, self.radius, rows=self.rows)
            stream = kd_tree._stream()
            if self.mode == "nearest":
                out = torch.empty((n_t, channels), dtype=data_dev.dtype, device="cuda")
                _cabi.check(L.prs_resample_nearest(source.handle, ctypes.byref(tg), float(self.radius),
                                                   kd_tree._ptr(data_dev), channels * data_dev.element_size(),
                                                   kd_tree._ptr(self.fill), kd_tree._ptr(out), stream))
            else:
                acc = torch.float64 if channels > 1 else torch.float32
                out = torch.empty((n_t, channels), dtype=acc, device="cuda")
                sig = (ctypes.c_double * channels)(*([self.cfg.get("sigma", 25000.0)] * channels))
                _cabi.check(L.prs_resample_gauss(source.handle, ctypes.byref(tg), int(self.k),
                                                 float(self.radius), kd_tree._ptr(data_dev), 0, channels, sig, 0.0,
                                                 kd_tree._ptr(out), None, None, None, stream))
        e2.record()
        if record:
            self.b_events.append((e0, e1))
            self.q_events.append((e1, e2))
        if want_info:  # reading the index description makes the host wait for the build: last step only
            self.info = {"n_valid": int(source.n_valid), "n_indexed": int(source.index.n_indexed),
                         "cells_per_side": int(source.index.cells_per_side), "index_bytes": int(source.index.bytes)}
        return out

    def out_itemsize(self):
        if self.mode != "nearest" and self.channels > 1:
            return 8
        return self.data_dev.element_size() if self.data_dev is not None else self.data_itemsize

    def timed(self, steps, warmup, sync_all, sampler=None):
        """W warm-up steps, then exactly K timed steps; returns (ms_per_step list pieces, launches)."""
        from pyresample_b200 import _cabi
        torch, L = self.torch, _cabi.lib()
        out = None
        for _ in range(warmup):
            out = self.step()  # held like in the timed loop: torch's allocator caches both output buffers
        sync_all()
        if sampler is not None:
            sampler.active = True
        launches0 = L.prs_launch_count()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(steps):
            out = self.step(record=True, want_info=(i == steps - 1))
        t_end.record()
        sync_all()
        del out
        launches = L.prs_launch_count() - launches0
        q_all = [a.elapsed_time(b) for a, b in self.q_events]
        b_all = [a.elapsed_time(b) for a, b in self.b_events]
        self.q_events, self.b_events = [], []
        return t_start.elapsed_time(t_end), q_all, b_all, launches

    def roofline(self, q_all, b_ms, peak, peak_kind, traffic):
        n_t_rank = self.rows[1] * self.W
        itemsize = self.data_dev.element_size() if self.data_dev is not None else self.data_itemsize
        q_ms = float(np.mean(q_all))
        b_alg = algorithmic_bytes(self.info["n_indexed"], n_t_rank, self.channels, 4, itemsize, self.out_itemsize())
        achieved = b_alg / (q_ms * 1e-3) / 1e9
        kernel = {"nearest": "prs::classify_kernel + prs::query_kernel (fused target transform + k-NN + gather)",
                  "gauss": "prs::classify_kernel + prs::query_kernel (fused target transform + k-NN + Gaussian "
                           "weighting, float64 accumulation)",
                  "custom": "prs::classify_kernel + prs::query_kernel (k-NN, neighbour info) + weight function on "
                            "the distances (torch elementwise) + prs::gather_weighted_kernel"}[self.mode]
        return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind, "kernel_ms": q_ms,
                "algorithmic_bytes_per_launch": int(b_alg), "index_build_ms": b_ms,
                "kernel_ms_min": float(np.min(q_all)), "kernel_ms_median": float(np.median(q_all)),
                "kernel_mpixel_per_s": n_t_rank / (q_ms * 1e-3) / 1e6}


def load_peak():
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return peak, "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"


def load_traffic(cid, world, scale):
    """dram__bytes_read + write per step of the query-path kernels from the committed ncu --set full capture."""
    if world != 1 or scale != 1.0:
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            entry = json.load(fh).get("config%d" % cid)
        return int(entry["dram_bytes_per_step"]) if entry else None
    except Exception:
        return None


def reduce_max(dist, torch, values):
    t = torch.tensor(values, device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configs index + 1 (default 2 = configs[1])")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--ref-rows", type=int, default=0,
                    help="target rows per step of the CPU arm (0: as many as fit PRS_REF_BUDGET_S, all if possible)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the C5 sub-record")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from pyresample_b200 import _cabi, distributed, geometry, kd_tree
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]  # these levels print a version banner on stdout; stdout carries ONE JSON line
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _cabi.lib()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every rank derives the (seeded) geometry; the source ARRAYS that are used come from rank 0 via NCCL
    wl = Resident(args.config, args.scale, rank, world)
    cfg, area, H, W = wl.cfg, wl.area, wl.H, wl.W
    channels, k, mode, radius = wl.channels, wl.k, wl.mode, wl.radius
    host, dev, names, rows = wl.host, wl.dev, wl.names, wl.rows
    sampler = ClockSampler(local_rank)
    sampler.start()
    elapsed_ms, q_all, b_all, launches = wl.timed(args.steps, args.warmup, sync_all, sampler)
    if not sampler.samples:  # a very short timed region: take one sample right at its end, still under load
        time.sleep(float(os.environ.get("PRS_BENCH_CLOCK_PERIOD", "0.05")) * 1.5)
    clocks = sampler.stop()
    if os.environ.get("PRS_BENCH_DEBUG"):
        print("query ms per step:", " ".join("%.3f" % v for v in q_all), file=sys.stderr)
        print("build ms per step:", " ".join("%.3f" % v for v in b_all), file=sys.stderr)
    q_ms, b_ms = float(np.mean(q_all)), float(np.mean(b_all))
    if world > 1:
        elapsed_ms, q_ms, b_ms = reduce_max(dist, torch, [elapsed_ms, q_ms, b_ms])
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
        d2h = n_t_total * channels * wl.out_itemsize()
        sigmas = [cfg.get("sigma", 25000.0)] * channels if channels > 1 else cfg.get("sigma", 25000.0)

        def host_source():
            if "source_area" in cfg:
                return cfg["source_area"]
            return geometry.SwathDefinition(host["lon"].numpy(), host["lat"].numpy())

        def e2e_step():
            if world == 1:
                if mode == "nearest":
                    return kd_tree.resample_nearest(host_source(), host["data"].numpy(), area, radius)
                if mode == "gauss":
                    return kd_tree.resample_gauss(host_source(), host["data"].numpy(), area, radius, sigmas,
                                                  neighbours=k)
                return kd_tree.resample_custom(host_source(), host["data"].numpy(), area, radius, wl.weight_funcs,
                                               neighbours=k)
            # N > 1, swath source: source and result live in shared page-locked host memory; every rank uploads its
            # share of the source over its own PCIe link (all-gather over NVLink) and writes its rows of the result
            # straight into the one assembled grid
            if shared is not None:
                return distributed.resample_nearest_shared(shared["lon"], shared["lat"], shared["data"], shared["out"],
                                                           area, radius)
            # otherwise: rank 0 uploads, NCCL broadcast, every rank resamples and reads back its own rows
            for n in names:
                if rank == 0:
                    dev[n].copy_(host[n], non_blocking=True)
            distributed.broadcast_source([dev[n] for n in names])
            src = cfg["source_area"] if "source_area" in cfg else geometry.SwathDefinition(dev["lon"], dev["lat"])
            if mode == "nearest":
                blk = distributed.resample_nearest_sharded(src, dev["data"], area, radius)
            else:
                funcs = wl.weight_funcs if mode == "custom" else (
                    [kd_tree._make_gauss(s) for s in sigmas] if channels > 1 else kd_tree._make_gauss(sigmas))
                blk = kd_tree._resample(src, dev["data"], area, 'custom', radius, neighbours=k, weight_funcs=funcs,
                                        _rows=rows)
            return kd_tree._host(blk)

        shared = None
        if world > 1 and mode == "nearest" and "source_area" not in cfg:
            shared = {n: distributed.share_host_array(host[n].numpy() if rank == 0 else None,
                                                      (wl.n_src, channels) if n == "data" else (wl.n_src,), cfg[n].dtype)
                      for n in names}
            shared["out"] = distributed.share_host_array(None, (H, W, channels), cfg["data"].dtype)
            h2d = sum(shared[n].nbytes for n in names)  # over all ranks: each uploads 1 / world of it
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
            e2e_ms = reduce_max(dist, torch, [e2e_ms])[0]
        if shared is not None:
            if rank == 0 and os.environ.get("PRS_BENCH_CHECK_SHARED"):
                want = kd_tree.resample_nearest(geometry.SwathDefinition(host["lon"].numpy(), host["lat"].numpy()),
                                                host["data"].numpy(), area, radius)
                assert np.array_equal(shared["out"].array.reshape(want.shape), want, equal_nan=True)
                print("shared result equals the single-GPU result", file=sys.stderr)
            dist.barrier()
            for sh in shared.values():
                sh.close()
        api = {"nearest": "resample_nearest", "gauss": "resample_gauss", "custom": "resample_custom"}[mode]
        e2e = {"value": n_t_total / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "host_memory": "pinned",
               "api": ("pyresample_b200.kd_tree.%s (numpy in, numpy out)" % api) if world == 1 else
                      ("pyresample_b200.distributed.resample_nearest_shared (source and assembled result in shared "
                       "page-locked host memory; every rank moves 1/N of each over its own PCIe link)"
                       if shared is not None else
                       "rank 0 uploads, NCCL broadcast, kd_tree resample of each rank's rows, read-back")}

    # ---- C4 as BASELINE.json words it: XArrayResamplerNN next to resample_custom (N = 1 only)
    xarray_nn = None
    if args.config == 4 and world == 1 and not args.no_e2e:
        xarray_nn = time_xarray_nn(cfg, radius, k)

    # ---- the configuration BASELINE.json names for the 8-GPU run (C5), resident, at this run's N
    extra = {}
    if args.config != 5 and not args.no_extra_configs and args.scale == 1.0:
        del wl.dev, dev
        data_itemsize = wl.data_dev.element_size()
        wl.data_dev = wl.source_dev = None
        wl.data_itemsize = data_itemsize
        torch.cuda.empty_cache()
        w5 = Resident(5, 1.0, rank, world)
        n5 = max(3, min(args.steps, 10))
        el5, q5, b5, _ = w5.timed(n5, 3, sync_all)
        q5m, b5m = float(np.mean(q5)), float(np.mean(b5))
        if world > 1:
            el5, q5m, b5m = reduce_max(dist, torch, [el5, q5m, b5m])
        peak, peak_kind = load_peak()
        extra["C5"] = {"config": workload_config(w5.cfg, 1.0), "n_gpus": world, "steps": n5, "warmup": 3,
                       "ms_per_step": el5 / n5, "value": w5.H * w5.W / (el5 / n5 * 1e-3) / 1e6, "unit": "Mpixel/s",
                       "dtype": arithmetic_dtype(w5.cfg), "query_ms": q5m, "index_build_ms": b5m,
                       "roofline": w5.roofline(q5, b5m, peak, peak_kind, load_traffic(5, world, 1.0)),
                       "index": dict(w5.info)}

    if rank == 0:
        peak, peak_kind = load_peak()
        roofline = wl.roofline(q_all, b_ms, peak, peak_kind, load_traffic(args.config, world, args.scale))
        if world > 1:
            roofline["kernel_ms_max_over_ranks"] = q_ms
        config = workload_config(cfg, args.scale)
        config["l2"] = ("inputs (%.0f MB index + data, %.0f MB output) exceed the 126 MB L2; no flush needed"
                        % ((wl.info["index_bytes"] + wl.n_src * channels * 4) / 1e6,
                           n_t_total * channels * wl.out_itemsize() / 1e6))
        config["sharding"] = ("target rows in stripes of %d dealt round-robin over %d rank(s); source broadcast once "
                              "(NCCL), each rank indexes only the source points near its rows"
                              % (distributed.DEFAULT_STRIPE, world))
        line = {"metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": arithmetic_dtype(cfg), "data": "synthetic", "config": config,
                "index": {"valid_source_points": wl.info["n_valid"], "indexed_points_rank0": wl.info["n_indexed"],
                          "cells_per_face_side": wl.info["cells_per_side"], "bytes": wl.info["index_bytes"]},
                "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e}
        if xarray_nn:
            line["xarray_nn"] = xarray_nn
        if extra:
            line["configs"] = extra
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(cfg, os.cpu_count() or 1)
            # one step of the CPU arm, bounded to ~20 s: all rows when they fit, else a sample scaled to the grid
            ref.set_rows(48)
            ref.step()
            per_row = ref.t_query / len(ref.rows)
            ref.set_rows(H if (20.0 - ref.t_build) >= per_row * H else max(16, (20.0 - ref.t_build) / per_row))
            v = float(ref.step())
            line["cpu_baseline"] = {"value": v, "unit": "Mpixel/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": ref.describe()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def time_xarray_nn(cfg, radius, k):
    """XArrayResamplerNN on C4: neighbours=k indices (float64 tree, as the reference's class builds it), then
    neighbours=1 indices + sampling of one channel (the class cannot sample k > 1, kd_tree.py:1082-1084)."""
    import torch

    from pyresample_b200 import geometry, kd_tree, xarray_nn
    if kd_tree.DataArray is None:  # xarray is not installed in this image: the class only needs these attributes
        class _DA(object):
            def __init__(self, data, dims=None, coords=None, attrs=None):
                self.data, self.dims, self.coords, self.attrs = data, tuple(dims), dict(coords or {}), dict(attrs or {})
            shape = property(lambda s: s.data.shape)
            ndim = property(lambda s: s.data.ndim)
            __getitem__ = lambda s, key: s.data[key]  # noqa: E731
            dtype = property(lambda s: s.data.dtype)
            sizes = property(lambda s: dict(zip(s.dims, s.data.shape)))
        kd_tree.DataArray = _DA
    DA = kd_tree.DataArray
    lon, lat = cfg["lon"], cfg["lat"]
    swath = geometry.SwathDefinition(DA(lon, dims=('y', 'x')), DA(lat, dims=('y', 'x')))
    area = cfg["area"]
    data2d = DA(np.ascontiguousarray(cfg["data"][:, 0]).reshape(lon.shape), dims=('y', 'x'))
    out = {}
    for name, nb in (("neighbours=%d get_neighbour_info (device search; index_array not read back)" % k, k),
                     ("neighbours=1 get_neighbour_info + get_sample_from_neighbour_info (numpy out)", 1)):
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = kd_tree.XArrayResamplerNN(swath, area, radius_of_influence=radius, neighbours=nb)
            r.get_neighbour_info()
            if nb == 1:
                res = r.get_sample_from_neighbour_info(data2d)
                np.asarray(res.data.compute())
            else:
                r._state.query()  # what computing index_array triggers, without the 8 GB int64 host copy
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
            del r
        out[name] = {"ms": float(np.min(ts)), "mpixel_per_s": area.shape[0] * area.shape[1] / (np.min(ts) * 1e-3) / 1e6}
    out["api"] = "pyresample_b200.kd_tree.XArrayResamplerNN (pageable numpy swath in; wall clock, best of 3)"
    _ = xarray_nn
    return out


if __name__ == "__main__":
    main()
