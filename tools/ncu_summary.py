"""Reduce `ncu -i X.ncu-rep --page raw --csv` to the per-kernel figures the profiles/ summaries quote.

usage: python tools/ncu_summary.py raw.csv [title]

For every kernel name: launches captured, mean time, DRAM bytes read / written per launch, L2 hit rate, issue-active,
achieved occupancy, registers, grid, and the warp-state (stall) breakdown in percent of sampled warp states.
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    i = col.get(name)
    if i is None:
        return None
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return None


def scaled(r, name, to):
    """Value of a metric converted to `to` (bytes: 'byte', time: 'us')."""
    v = num(r, name)
    if v is None:
        return None
    u = units[col[name]].lower()
    f = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3,
         "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    return v * f


stall_cols = [h for h in hdr if re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", h) and "not_issued" not in h]
by = collections.OrderedDict()
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    by.setdefault(r[col["Kernel Name"]], []).append(r)

print("# %s" % title)
print("# from ncu --set full --clock-control none (cold caches, kernels serialised): shares, not absolute times")
for name, rs in by.items():
    def mean(metric, conv=None):
        vals = [scaled(r, metric, conv) if conv else num(r, metric) for r in rs]
        vals = [v for v in vals if v is not None]
        return sum(vals) / len(vals) if vals else float("nan")
    short = re.sub(r"\(.*", "", name)
    print("\nkernel %s   [%d launch(es)]" % (short, len(rs)))
    print("  time %.1f us   grid %d x %d thr   regs %d   dyn smem %.0f B" % (
        mean("gpu__time_duration.sum", "us"), mean("launch__grid_size"), mean("launch__block_size"),
        mean("launch__registers_per_thread"), mean("launch__shared_mem_per_block_dynamic", "byte")))
    rd, wr = mean("dram__bytes_read.sum", "byte"), mean("dram__bytes_write.sum", "byte")
    print("  dram read %.1f MB  write %.1f MB  (sum %.1f MB)   = %.0f GB/s over the kernel's (cold) duration" % (
        rd / 1e6, wr / 1e6, (rd + wr) / 1e6, (rd + wr) / 1e3 / max(mean("gpu__time_duration.sum", "us"), 1e-9)))
    print("  L2 hit %.1f %%   L1 hit %.1f %%   issue active %.1f %%   achieved occupancy %.1f %%   threads/inst %.1f" % (
        mean("lts__t_sector_hit_rate.pct"), mean("l1tex__t_sector_hit_rate.pct"),
        mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        mean("sm__warps_active.avg.pct_of_peak_sustained_active"),
        mean("smsp__thread_inst_executed_per_inst_executed.ratio")))
    st = {}
    for h in stall_cols:
        st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = sum(num(r, h) or 0 for r in rs)
    tot = sum(st.values())
    if tot:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:8]
        print("  warp states: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in top))
