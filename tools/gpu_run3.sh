ncu --kernel-name 'regex:(qmf_analysis|mdct|alloc|quant_pack|unpack|synth)_kernel' --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 18 --csv --log-file gpurun_out/traffic_v38.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/traffic_v38.log 2>&1
tail -c 600 gpurun_out/traffic_v38.csv
