#!/bin/bash
# Development aid: role kernels side by side (fork / join) against one stream, on the closure calls and config 4.
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_fork.log 2>&1; tail -2 gpurun_out/pytest_fork.log
python tools/closure_latency.py > gpurun_out/closure_fork.txt 2>&1
CARTA1_NO_FORK=1 python tools/closure_latency.py > gpurun_out/closure_nofork.txt 2>&1
head -9 gpurun_out/closure_fork.txt; echo; head -9 gpurun_out/closure_nofork.txt
python tools/bench_cfg4.py > gpurun_out/cfg4_fork.txt 2>&1; CARTA1_NO_FORK=1 python tools/bench_cfg4.py > gpurun_out/cfg4_nofork.txt 2>&1
tail -8 gpurun_out/cfg4_fork.txt; echo; tail -8 gpurun_out/cfg4_nofork.txt
