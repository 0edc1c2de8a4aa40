#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
./tools/micro/sector_probe > $O/probe_plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__time_duration.sum --csv --log-file $O/probe.csv ./tools/micro/sector_probe > $O/probe.log 2>&1
grep -E "probe" $O/probe.csv | awk -F'","' '{gsub(/"/,"",$NF); print $(NF-2), $NF}'
