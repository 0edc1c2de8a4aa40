#!/bin/bash
# tools/gpu_long_variants.sh tag [tag ...] : long-block sweep (mu 8.5 and 6.5, records cut at 4096 bases) for prebuilt library variants
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for tag in "$@"; do
  if [ "$tag" == "main" ]; then unset GAT_LIB_PATH; else export GAT_LIB_PATH=$PWD/genomealignmenttools_b200/_build/libgat_$tag.so; fi
  for mu in 8.5 6.5; do
    extra="--split 4096 --blocks 2000000"; [ "$mu" == "6.5" ] && extra="--split 4096 --blocks 4000000"
    timeout 900 python bench.py --no-cpu-baseline --steps 20 --warmup 3 --mean-log-len $mu $extra > $O/bench_${tag}_mu$mu.json 2> $O/bench_${tag}_mu$mu.err
    python -c "
import json
d=json.loads(open('$O/bench_${tag}_mu$mu.json').read().strip().splitlines()[-1])
print('$tag mu $mu', d['roofline']['kernel'], 'kernel_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'parity',d.get('parity_mismatches', d.get('parity')))"
  done
done
