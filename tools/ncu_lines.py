#!/usr/bin/env python
"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo + --import-source on):

    python tools/ncu_lines.py report.ncu-rep query_kernel [top_n]

Prints the source lines that execute the most warp instructions / collect the most stall samples, with the
average number of active threads, from `ncu --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys


def function_starts(fname):
    """[(first line, name)] of the device functions defined in pyresample_b200/csrc/<fname> (by their signature lines)."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyresample_b200", "csrc", fname)
    starts = []
    try:
        src = open(path).read().splitlines()
    except OSError:
        return starts
    sig = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__)[^;(]*?\b([A-Za-z_]\w*)\s*\($")
    for n, line in enumerate(src, 1):
        if not line or line[0] in " \t/#}":
            continue
        head = line.split("(")[0] + "("
        m = sig.match(head)
        if m:
            starts.append((n, m.group(1)))
        elif n >= 2 and src[n - 2].startswith("template") and re.match(r"^\s+\w+\(", line):
            pass
    # a signature split over two lines: `__global__ void __launch_bounds__(..)\n    name(`
    for n, line in enumerate(src, 1):
        m = re.match(r"^    ([a-z_]\w*)\(const ", line)
        if m and n >= 2 and ("__global__" in src[n - 2] or "__device__" in src[n - 2]):
            starts.append((n - 1, m.group(1)))
    return sorted(set(starts))


def function_of(starts, line):
    name = "?"
    for n, f in starts:
        if n > line:
            break
        name = f
    return name


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines = []
    cur_file, hdr = None, None
    seen_kernel = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            if seen_kernel is None:
                seen_kernel = r[1]
            elif r[1] != seen_kernel:
                cur_file = None  # a second kernel instance matched: keep the first only
        elif r[0] == "Line No":
            hdr = r
        elif cur_file and hdr and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            try:
                inst = int(d["Instructions Executed"])
                tinst = int(d["Thread Instructions Executed"])
                samp = int(d["Warp Stall Sampling (All Samples)"])
            except (KeyError, ValueError):
                continue
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k
                      and v.isdigit() and int(v)}
            lines.append((cur_file, int(r[0]), r[1].strip()[:90], inst, tinst, samp, stalls))
    tot_i = sum(l[3] for l in lines) or 1
    tot_s = sum(l[5] for l in lines) or 1
    print("kernel:", seen_kernel)
    print("total warp instructions %d, thread instructions %d (%.1f active threads / instruction), samples %d"
          % (tot_i, sum(l[4] for l in lines), sum(l[4] for l in lines) / tot_i, tot_s))
    agg = {}
    for l in lines:
        for k, v in l[6].items():
            agg[k] = agg.get(k, 0) + v
    print("stall samples:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_s) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    by_fn, cache = {}, {}
    for l in lines:
        if l[0] not in cache:
            cache[l[0]] = function_starts(l[0])
        key = "%s:%s" % (l[0], function_of(cache[l[0]], l[1])) if cache[l[0]] else l[0]
        a = by_fn.setdefault(key, [0, 0, 0, 0])
        a[0] += l[3]
        a[1] += l[4]
        a[2] += l[5]
        a[3] += l[6].get("long_sb", 0)
    print("\n-- by function (source lines attributed to the function whose definition precedes them)")
    for key, a in sorted(by_fn.items(), key=lambda kv: -kv[1][2])[:24]:
        print("%5.1f%% samp %5.1f%% inst  thr %4.1f  long_sb %4.1f%%  %s" % (100.0 * a[2] / tot_s, 100.0 * a[0] / tot_i,
                                                                         a[1] / max(a[0], 1), 100.0 * a[3] / tot_s, key))
    print("\n-- by warp instructions executed")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print("%5.1f%% inst %5.1f%% samp  thr %4.1f  %s:%d  %s" % (100.0 * l[3] / tot_i, 100.0 * l[5] / tot_s,
                                                               l[4] / max(l[3], 1), l[0], l[1], l[2]))
    print("\n-- by stall samples")
    for l in sorted(lines, key=lambda l: -l[5])[:top]:
        top_st = ", ".join("%s %d" % kv for kv in sorted(l[6].items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%% samp %5.1f%% inst  %s:%d  %s   [%s]" % (100.0 * l[5] / tot_s, 100.0 * l[3] / tot_i, l[0], l[1],
                                                            l[2], top_st))


if __name__ == "__main__":
    main()
