#!/bin/bash
# tools/gpu_ncu.sh [tag] : ncu full capture of the scoring kernel of one library variant (after a plain run of the same command)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
tag=${1:-main}
if [ "$tag" != "main" ]; then export GAT_LIB_PATH=$PWD/genomealignmenttools_b200/_build/libgat_$tag.so; fi
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/plain_$tag.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scoreTiles -s 3 -c 1 -o $O/prof_$tag -f python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/ncu_$tag.log 2>&1
echo "ncu $tag rc=$?"
