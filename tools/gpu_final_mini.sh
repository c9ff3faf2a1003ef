#!/bin/bash
# Refresh of the round-end pass after a change that touches C5 / the default line only: -m gpu suite, default bench line,
# C5 and C1 bench lines, C5 launch list and ncu --set full summary.
set -u
O=gpurun_out/${1:-final3}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 rc=$?" >> $O/status.txt
for c in 5 1; do
  timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench_c$c.json 2> $O/bench_c$c.err; echo "bench c$c rc=$?" >> $O/status.txt
done
T=/tmp/ncu_tmp; mkdir -p $T
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c5.csv \
    python bench.py --config 5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $T/ncu_list_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'classify_kernel|query_kernel|area_lonlats' --launch-skip 3 -c 3 \
    -o $T/full_c5 -f python bench.py --config 5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $T/ncu_full_c5.log 2>&1
ncu -i $T/full_c5.ncu-rep --page raw --csv > $T/full_c5_raw.csv 2>/dev/null
python tools/ncu_summary.py $T/full_c5_raw.csv "ncu --set full, bench.py --config 5 --steps 2 --warmup 1 (kernels: area_lonlats, classify_kernel, query_kernel)" > $O/ncu_c5.txt 2>&1
for kern in classify_kernel query_kernel; do python tools/ncu_lines.py $T/full_c5.ncu-rep $kern 30 > $O/ncu_c5_lines_$kern.txt 2>&1; done
cat $O/status.txt; tail -3 $O/pytest.log
