python tools/closure_latency.py > gpurun_out/closure_now.txt 2>&1
CARTA1_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/closure_launches.csv python tools/one_call_probe.py > gpurun_out/closure_ncu.log 2>&1
tail -3 gpurun_out/closure_now.txt
