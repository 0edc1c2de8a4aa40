#!/bin/bash
# tools/measure_traffic.sh -- on the GPU box: DRAM bytes of one launch of the scoring kernel on the default bench workload
# (N = 1, 10 M blocks), measured by ncu after the same command ran clean without it.  Writes gpurun_out/traffic.json;
# copy it to profiles/traffic.json (bench.py quotes it as roofline.traffic only while the kernel sources still hash to
# the value recorded here).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/traffic_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:scoreTiles -s 3 -c 1 \
    --csv --log-file $O/traffic_ncu.csv python bench.py --no-cpu-baseline --steps 2 --warmup 3 > $O/traffic_ncu.log 2>&1 || { echo "ncu run failed"; exit 1; }
python - <<'PY'
import csv, json, sys
sys.path.insert(0, '.')
import bench
vals = {}
for row in csv.reader(open('gpurun_out/traffic_ncu.csv')):
    if len(row) > 3 and row[-3].startswith(('dram__', 'gpu__')):
        vals[row[-3]] = float(row[-1].replace(',', ''))
out = {"dram_bytes_per_launch": int(vals['dram__bytes_read.sum'] + vals['dram__bytes_write.sum']),
       "dram_bytes_read": int(vals['dram__bytes_read.sum']), "dram_bytes_write": int(vals['dram__bytes_write.sum']),
       "kernel": "scoreTilesKernel<true, true, false>", "kernel_us_under_ncu": vals['gpu__time_duration.sum'] / 1e3,
       "blocks": 10000000, "source_hash": bench.source_hash(),
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:scoreTiles -s 3 -c 1, python bench.py --no-cpu-baseline --steps 2 --warmup 3 (tools/measure_traffic.sh)"}
json.dump(out, open('gpurun_out/traffic.json', 'w'), indent=1)
print(out)
PY
