#!/bin/bash
# full GPU test suite on the in-tree library, then bench lines of the given variants
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_full.log
[ $# -gt 0 ] && bash tools/gpu_variants.sh "$@"
