#!/bin/bash
# Round-end GPU pass: the whole -m gpu suite, every bench line, launch lists and ncu --set full captures reduced ON THE
# BOX to text summaries (the .ncu-rep files stay in /tmp there: gpurun_out/ must remain under 64 MiB).
set -u
O=gpurun_out/${1:-final}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 --durations=12 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 rc=$?" >> $O/status.txt
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_c2_reference_arm.json 2> $O/bench_c2_reference_arm.err; echo "reference arm rc=$?" >> $O/status.txt
for c in 3 4 5 1; do
  timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench_c$c.json 2> $O/bench_c$c.err; echo "bench c$c rc=$?" >> $O/status.txt
done
timeout 600 python tools/time_api.py 2 > $O/api_timings_c2.txt 2>&1; echo "api timings rc=$?" >> $O/status.txt
T=/tmp/ncu_tmp; mkdir -p $T
for c in 2 3 4 5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c$c.csv \
      python bench.py --config $c --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $T/ncu_list_c$c.log 2>&1
  echo "launch list c$c rc=$?" >> $O/status.txt
done
full() {  # name, config, kernel regex, launch-skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip $4 -c $5 \
      -o $T/full_$1 -f python bench.py --config $2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $T/ncu_full_$1.log 2>&1
  echo "ncu full $1 rc=$?" >> $O/status.txt
  ncu -i $T/full_$1.ncu-rep --page raw --csv > $T/full_$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py $T/full_$1_raw.csv "ncu --set full, bench.py --config $2 --steps 2 --warmup 1 (kernels: $3)" > $O/ncu_$1.txt 2>&1
  for kern in classify_kernel query_kernel gather_weighted build_select build_xyz build_scatter; do
    if grep -q "$kern" $T/full_$1_raw.csv; then python tools/ncu_lines.py $T/full_$1.ncu-rep $kern 30 > $O/ncu_$1_lines_$kern.txt 2>&1; fi
  done
}
full c2 2 'classify_kernel|query_kernel' 2 2
full c2_build 2 'build_|scan_' 12 12
full c3 3 'classify_kernel|query_kernel' 2 2
full c4 4 'classify_kernel|query_kernel|gather_weighted|weight_distances' 4 4
full c5 5 'classify_kernel|query_kernel' 2 2
cat $O/status.txt
tail -5 $O/pytest.log
for c in 2 3 4 5 1; do tail -c 400 $O/bench_c$c.err; done
du -sh gpurun_out
