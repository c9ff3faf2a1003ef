#!/bin/bash
# First GPU pass of round 2: the whole -m gpu suite, bench lines of C2..C5, launch lists and ncu --set full
# captures of the query kernels of C3 / C4 / C5 (each only after its command ran clean without ncu).
set -u
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/smi.txt 2>&1
python -m pytest tests -m gpu -x -q --durations=12 -s > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 rc=$?" >> $O/status.txt
for c in 3 4 5 1; do
  python bench.py --config $c --steps 5 --warmup 3 --no-extra-configs > $O/bench_c$c.json 2> $O/bench_c$c.err; echo "bench c$c rc=$?" >> $O/status.txt
done
for c in 3 4 5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c$c.csv \
      python bench.py --config $c --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_list_c$c.log 2>&1
  echo "launch list c$c rc=$?" >> $O/status.txt
done
# one full capture per configuration: the second step's classify + query (+ weighted gather) kernels
ncu --set full --clock-control none --import-source on -k regex:'classify_kernel|query_kernel' --launch-skip 2 -c 2 \
    -o $O/full_c3 -f python bench.py --config 3 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_full_c3.log 2>&1
echo "ncu full c3 rc=$?" >> $O/status.txt
ncu --set full --clock-control none --import-source on -k regex:'classify_kernel|query_kernel|gather_weighted' --launch-skip 3 -c 3 \
    -o $O/full_c4 -f python bench.py --config 4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_full_c4.log 2>&1
echo "ncu full c4 rc=$?" >> $O/status.txt
ncu --set full --clock-control none --import-source on -k regex:'classify_kernel|query_kernel' --launch-skip 2 -c 2 \
    -o $O/full_c5 -f python bench.py --config 5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_full_c5.log 2>&1
echo "ncu full c5 rc=$?" >> $O/status.txt
ncu --set full --clock-control none --import-source on -k regex:'build_|scan_' --launch-skip 12 -c 12 \
    -o $O/full_c2_build -f python bench.py --config 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-configs > $O/ncu_full_c2b.log 2>&1
echo "ncu full c2 build rc=$?" >> $O/status.txt
cat $O/status.txt
tail -3 $O/pytest.log
