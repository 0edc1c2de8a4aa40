#!/bin/bash
# One GPU call: smoke, parity tests, bench (new kernel, old kernel), ncu launch list + full capture of the scoring kernel.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 $O/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > $O/bench_new.json 2> $O/bench_new.err; echo "bench new rc=$?"; tail -c 1500 $O/bench_new.json
GAT_KERNEL=chunks timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > $O/bench_old.json 2> $O/bench_old.err; echo "bench old rc=$?"; tail -c 600 $O/bench_old.json
if [ "$1" == "ncu" ]; then
  timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/ncu1.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scoreTiles -s 3 -c 1 -o $O/prof_tiles python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/ncu2.log 2>&1
  echo "ncu full rc=$?"
fi
