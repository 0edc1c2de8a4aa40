#!/bin/bash
# tools/gpu_variants.sh tag [tag ...] : on the GPU box, quick parity (smoke + the parity test file) and the bench for each
# prebuilt variant genomealignmenttools_b200/_build/libgat_<tag>.so (built here with tools/variants.sh); "main" = the in-tree library.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for tag in "$@"; do
  if [ "$tag" == "main" ]; then unset GAT_LIB_PATH; else export GAT_LIB_PATH=$PWD/genomealignmenttools_b200/_build/libgat_$tag.so; fi
  timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke_$tag.log 2>&1; rc=$?
  if [ $rc -ne 0 ]; then echo "$tag: smoke FAILED rc=$rc"; tail -3 $O/smoke_$tag.log; continue; fi
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/parity_$tag.log 2>&1; prc=$?
  timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 3 ${BENCH_ARGS} > $O/bench_$tag.json 2> $O/bench_$tag.err; brc=$?
  python - "$tag" "$prc" "$brc" <<'PY'
import json,sys
tag,prc,brc=sys.argv[1:4]
try:
    d=json.loads(open('gpurun_out/bench_%s.json'%tag).read().strip().splitlines()[-1])
    print(tag,'parity_rc',prc,'kernel_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'step_ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1))
except Exception as e:
    print(tag,'parity_rc',prc,'bench_rc',brc,'no bench line',e)
PY
done
