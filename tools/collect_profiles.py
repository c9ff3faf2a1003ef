"""Copy the summaries of a round-end GPU pass (tools/gpu_final.sh -> gpurun_out/<dir>) into profiles/ under the
round's names, and refresh profiles/traffic.json from the ncu --set full summaries.

    python tools/collect_profiles.py gpurun_out/final r2
"""
import json
import os
import re
import shutil
import sys

src, tag = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(root, "profiles")


def copy(name, to):
    a = os.path.join(src, name)
    if os.path.exists(a) and os.path.getsize(a):
        shutil.copyfile(a, os.path.join(dst, "%s_%s" % (tag, to)))
        return True
    return False


for c in (1, 2, 3, 4, 5):
    copy("bench_c%d.json" % c, "bench_c%d.json" % c)
copy("bench_c2_reference_arm.json", "bench_c2_reference_arm.json")
copy("api_timings_c2.txt", "api_timings_c2.txt")
copy("pytest.log", "pytest_gpu.log")
for c in (2, 3, 4, 5):
    copy("launches_c%d.csv" % c, "launches_c%d.csv" % c)
for name in sorted(os.listdir(src)):
    if name.startswith("ncu_") and name.endswith(".txt"):
        copy(name, name)

# traffic.json: dram bytes per step of the query-path kernels (classification + search [+ weighted gather])
traffic = {}
for c in (2, 3, 4, 5):
    path = os.path.join(src, "ncu_c%d.txt" % c)
    if not os.path.exists(path):
        continue
    total, kernels = 0.0, []
    cur = None
    for line in open(path):
        m = re.match(r"kernel (?:void )?(\S+)", line)
        if m:
            cur = m.group(1)
        m = re.search(r"\(sum ([0-9.]+) MB\)", line)
        if m and cur and re.search(r"classify_kernel|query_kernel|gather_weighted|weight_distances", cur):
            total += float(m.group(1)) * 1e6
            kernels.append(cur)
    if kernels:
        traffic["config%d" % c] = {
            "scale": 1.0, "n_gpus": 1, "kernels": " + ".join(kernels), "dram_bytes_per_step": int(total),
            "source": "profiles/%s_ncu_c%d.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
                      % (tag, c)}
if traffic:
    with open(os.path.join(dst, "traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)
        fh.write("\n")
print("traffic:", {k: v["dram_bytes_per_step"] for k, v in traffic.items()})
