#!/bin/bash
# Quick GPU pass: the -m gpu suite and the resident step of the configurations named on the command line.
set -u
O=gpurun_out/${1:-q}
shift
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --maxfail=4 -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
tail -3 $O/pytest.log
for c in "$@"; do
  timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-extra-configs --no-e2e --no-cpu-baseline > $O/bench_c$c.json 2> $O/bench_c$c.err; echo "bench c$c rc=$?" >> $O/status.txt
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_c$c.json")); r=d["roofline"]
    print("C$c ms/step %.3f query %.3f build %.3f frac %.3f launches %d" % (d["ms_per_step"], r["kernel_ms"], r["index_build_ms"], r["frac"], d["gpu_launches"]))
except Exception as e:
    print("C$c failed", e); print(open("$O/bench_c$c.err").read()[-1500:])
PY
done
T=/tmp/ncu_tmp; mkdir -p $T
if [ -n "${PRS_NCU:-}" ]; then
  for c in $PRS_NCU; do
    ncu --set full --clock-control none --import-source on -k regex:'classify_kernel|query_kernel|gather_weighted' --launch-skip 3 -c 3 \
      -o $T/full_c$c -f python bench.py --config $c --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $T/ncu_c$c.log 2>&1
    ncu -i $T/full_c$c.ncu-rep --page raw --csv > $O/full_c${c}_raw.csv 2>/dev/null
    for kern in classify_kernel query_kernel gather_weighted; do
      if grep -q "$kern" $O/full_c${c}_raw.csv; then python tools/ncu_lines.py $T/full_c$c.ncu-rep $kern 40 > $O/full_c${c}_lines_$kern.txt 2>&1; fi
    done
  done
fi
cat $O/status.txt
