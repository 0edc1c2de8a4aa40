cd /root/repo
for b in 1250000 5000000; do BENCH_ARGS="--blocks $b" bash tools/gpu_traffic.sh main; mv gpurun_out/traffic_main.csv gpurun_out/traffic_main_$b.csv; done
