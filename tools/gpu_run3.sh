python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v38.json 2> gpurun_out/bench_v38.err; tail -c 1500 gpurun_out/bench_v38.json
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/traffic_v38.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/traffic_v38.log 2>&1
tail -c 300 gpurun_out/traffic_v38.csv
