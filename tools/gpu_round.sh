#!/bin/bash
# tools/gpu_round.sh [traffic|launches|full] : what the driver runs at round end (full GPU tests, smoke, bench, reference arm),
# then one ncu pass (one per call)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
if [ "$1" != "launches" ] && [ "$1" != "full" ]; then
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_full.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_r02.json 2> $O/bench_r02.err; echo "bench rc=$?"; tail -c 2500 $O/bench_r02.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r02_reference.json 2> $O/bench_r02_reference.err; echo "ref rc=$?"; tail -c 600 $O/bench_r02_reference.json
fi
case "$1" in
  traffic) bash tools/measure_traffic.sh ;;
  launches) timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/plain_launch.log 2>&1 &&
            timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/ncu_launch.log 2>&1; echo "launch list rc=$?" ;;
  full) bash tools/gpu_ncu.sh main ;;
esac
