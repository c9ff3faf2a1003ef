#!/bin/bash
# per-function profile of the search kernels of the configurations named on the command line
set -u
O=gpurun_out/fn; mkdir -p $O; T=/tmp/ncu_tmp; mkdir -p $T
for c in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:'query_kernel' --launch-skip 2 -c 1 \
      -o $T/fn_c$c -f python bench.py --config $c --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $T/fn_c$c.log 2>&1
  python tools/ncu_lines.py $T/fn_c$c.ncu-rep query_kernel 25 > $O/c${c}_query_kernel.txt 2>&1
  sed -n 1,32p $O/c${c}_query_kernel.txt | cut -c1-150
done
