#!/bin/bash
# tools/gpu_scale.sh N [extra bench args]: bench.py on N GPUs of the box (strong scaling by default) -> gpurun_out/scale_N.json
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=$1; shift
if [ "$N" == "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline "$@" > $O/scale_$N.json 2> $O/scale_$N.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline "$@" > $O/scale_$N.json 2> $O/scale_$N.err
fi
echo "N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], d['scaling'], 'value', round(d['value'],1), 'step_ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'ceil', d['e2e'].get('h2d_ceiling_gbs_all_ranks'), d.get('scaling_check'))
except Exception as e:
    print('no line', e)
PY
tail -5 $O/scale_$N.err
