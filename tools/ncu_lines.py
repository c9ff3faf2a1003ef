#!/usr/bin/env python
"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo + --import-source on):

    python tools/ncu_lines.py report.ncu-rep query_kernel [top_n]

Prints the source lines that execute the most warp instructions / collect the most stall samples, with the
average number of active threads, from `ncu --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys


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
