#!/bin/bash
# kernel experiments: resident step of C3 / C4 under different cell occupancies
set -u
O=gpurun_out/${1:-exp}
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_baseline_configs.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for lam in default 1 2 3; do
  for c in 3 4; do
    if [ $lam = default ]; then unset PRS_CELL_OCCUPANCY; else export PRS_CELL_OCCUPANCY=$lam; fi
    timeout 600 python bench.py --config $c --steps 5 --warmup 2 --no-extra-configs --no-e2e --no-cpu-baseline > $O/b.json 2> $O/b.err
    python - <<PY
import json
try:
    d=json.load(open("$O/b.json")); r=d["roofline"]
    print("lambda $lam C$c ms/step %.3f query %.3f build %.3f cells/side %s" % (d["ms_per_step"], r["kernel_ms"], r["index_build_ms"], d["index"]["cells_per_face_side"]))
except Exception as e:
    print("lambda $lam C$c failed", e); print(open("$O/b.err").read()[-800:])
PY
  done
done
