#!/bin/bash
set -u
O=gpurun_out/${1:-var}
mkdir -p $O
for v in b200 v1 v2 v3 v4 v5; do
  export PRS_B200_LIBRARY=$PWD/pyresample_b200/libprs_$v.so
  for c in 3 4; do
    timeout 600 python bench.py --config $c --steps 5 --warmup 2 --no-extra-configs --no-e2e --no-cpu-baseline > $O/b.json 2> $O/b.err
    python - <<PY
import json
try:
    d=json.load(open("$O/b.json")); r=d["roofline"]
    print("$v C$c ms/step %.3f query %.3f build %.3f" % (d["ms_per_step"], r["kernel_ms"], r["index_build_ms"]))
except Exception as e:
    print("$v C$c failed", e); print(open("$O/b.err").read()[-600:])
PY
  done
done
