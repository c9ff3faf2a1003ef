#!/bin/bash
# A/B of library variants (tools/build_variants.py) on the C3 / C4 search kernels; the product library also runs the
# -m gpu suite and the other configurations named after the output directory.
set -u
O=gpurun_out/${1:-var}
shift
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --maxfail=4 -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for v in b200 v1 v2 v3 v4 v5; do
  [ -f pyresample_b200/libprs_$v.so ] || continue
  export PRS_B200_LIBRARY=$PWD/pyresample_b200/libprs_$v.so
  cfgs="3 4"; [ $v = b200 ] && cfgs="3 4 $*"
  for c in $cfgs; do
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
