#!/bin/bash
# Build kernel variants of libgat.so into genomealignmenttools_b200/_build/ :  tools/variants.sh tag "-DGAT_MIN_CTAS=4 ..." [tag flags ...]
# Run them on the GPU box with:  tools/variants.sh --run tag [tag ...]   (prints kernel ms per variant)
set -e
cd "$(dirname "$0")/.."
OUT=genomealignmenttools_b200/_build
mkdir -p $OUT
if [ "$1" == "--run" ]; then
  shift
  for tag in "$@"; do
    GAT_LIB_PATH=$PWD/$OUT/libgat_$tag.so python bench.py --no-cpu-baseline --steps 20 --warmup 3 ${BENCH_ARGS} 2>/dev/null |
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'step_ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"
  done
  exit 0
fi
while [ $# -gt 1 ]; do
  tag=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Igenomealignmenttools_b200/csrc \
     $flags -Xptxas -v --shared -o $OUT/libgat_$tag.so genomealignmenttools_b200/csrc/gat_capi.cu 2>&1 | grep -A2 "scoreTilesKernelILb1ELb1ELb1" | grep -E "registers|spill" | sed "s/^/$tag: /" &
done
wait
