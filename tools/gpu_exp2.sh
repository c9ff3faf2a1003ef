#!/bin/bash
set -u
O=gpurun_out/${1:-exp2}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest.log | cut -c1-200
run() {  # label config env...
  timeout 600 python bench.py --config $2 --steps 5 --warmup 2 --no-extra-configs --no-e2e --no-cpu-baseline > $O/b.json 2> $O/b.err
  python - <<PY
import json
try:
    d=json.load(open("$O/b.json")); r=d["roofline"]
    print("$1 C$2 ms/step %.3f query %.3f build %.3f cells/side %s" % (d["ms_per_step"], r["kernel_ms"], r["index_build_ms"], d["index"]["cells_per_face_side"]))
except Exception as e:
    print("$1 C$2 failed", e); print(open("$O/b.err").read()[-600:])
PY
}
for lam in 6 8 12; do PRS_CELL_OCCUPANCY=$lam run "lambda=$lam" 3; done
for lam in 12 16 24; do PRS_CELL_OCCUPANCY=$lam run "lambda=$lam" 4; done
for lam in 3 4; do PRS_CELL_OCCUPANCY=$lam run "lambda=$lam" 5; done
for lam in 3 4; do PRS_CELL_OCCUPANCY=$lam run "lambda=$lam" 2; done
unset PRS_CELL_OCCUPANCY
run default 2; run default 5; run default 1
