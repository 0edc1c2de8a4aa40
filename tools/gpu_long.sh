#!/bin/bash
# block-length sweep: kernel-only roofline fraction for long blocks (records cut at 4096 bases)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/parity_long.log 2>&1; echo "parity rc=$?"; tail -3 $O/parity_long.log
for mu in 3.6 6.5 8.5; do
  extra=""; [ "$mu" == "6.5" ] && extra="--split 4096 --blocks 4000000"; [ "$mu" == "8.5" ] && extra="--split 4096 --blocks 2000000"
  timeout 900 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --mean-log-len $mu $extra > $O/bench_mu$mu.json 2> $O/bench_mu$mu.err
  python -c "
import json
d=json.loads(open('$O/bench_mu$mu.json').read().strip().splitlines()[-1])
print('mu $mu', d['roofline']['kernel'], 'kernel_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'step_ms',round(d['ms_per_step'],4),'Gbp/s',round(d['value'],1),'aligned Mbp',d['aligned_bp_per_gpu']/1e6)"
done
