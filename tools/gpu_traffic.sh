#!/bin/bash
# tools/gpu_traffic.sh tag [tag ...]: bench timing + DRAM / L2 / L1 sector counts of the scoring kernel (ncu, a handful of metrics) per variant
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
for tag in "$@"; do
  if [ "$tag" == "main" ]; then unset GAT_LIB_PATH; else export GAT_LIB_PATH=$PWD/genomealignmenttools_b200/_build/libgat_$tag.so; fi
  timeout 600 python bench.py --no-cpu-baseline --steps 20 --warmup 3 ${BENCH_ARGS} > $O/bench_$tag.json 2> $O/bench_$tag.err || { echo "$tag bench failed"; continue; }
  python -c "import json; d=json.loads(open('$O/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag','kernel_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'step_ms',round(d['ms_per_step'],4))"
  timeout 600 ncu --metrics $M --clock-control none -k regex:scoreTiles -s 3 -c 1 --csv --log-file $O/traffic_$tag.csv python bench.py --no-cpu-baseline --steps 2 --warmup 3 ${BENCH_ARGS} > $O/traffic_$tag.log 2>&1
  grep -E "dram__bytes|lookup_miss|op_read.sum|global_op_ld.sum|inst_executed|wavefronts|issue_active" $O/traffic_$tag.csv | awk -F'","' '{gsub(/"/,"",$NF); printf "   %s %s\n", $(NF-2), $NF}'
done
