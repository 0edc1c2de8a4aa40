#!/bin/bash
# measured lines for configs 1-4 of BASELINE.json (tools/bench_configs.py) -> gpurun_out/config_<k>.json
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for k in "$@"; do
  timeout 1500 python bench.py --config $k --steps 20 --warmup 3 > $O/config_$k.json 2> $O/config_$k.err; echo "config $k rc=$?"
  tail -3 $O/config_$k.err
  python -c "
import json
try:
    d=json.loads(open('$O/config_$k.json').read().strip().splitlines()[-1])
    print('  value',round(d['value'],1),'Gbp/s step_ms',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1),'cpu',d['cpu_baseline'].get('value'),'parity',d['parity']['mismatches'])
except Exception as e: print('  no line',e)
"
done
