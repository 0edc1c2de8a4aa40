#!/bin/bash
# A/B of programmatic dependent launch: parity, then the bench step with and without it at 10 M and 1.25 M blocks
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_pdl.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_pdl.log
for blocks in 10000000 1250000; do
  for pdl in 0 1; do
    GAT_NO_PDL=$pdl timeout 600 python bench.py --no-cpu-baseline --steps 50 --warmup 5 --blocks $blocks > $O/bench_pdl${pdl}_$blocks.json 2> $O/bench_pdl${pdl}_$blocks.err
    python -c "
import json
d=json.loads(open('$O/bench_pdl${pdl}_$blocks.json').read().strip().splitlines()[-1])
print('blocks $blocks no_pdl=$pdl step_ms',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),'value',round(d['value'],1))"
  done
done
