#!/bin/bash
# Development aid: what a round-end check runs on the GPU box, in one gpurun call.
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh TAG'
# Writes gpurun_out/{pytest,bench,bench_ref,launches}_TAG.*; copy what should be judged into profiles/
# (python tools/traffic_from_launches.py gpurun_out/launches_TAG.csv profiles/r02_dram_traffic_TAG.json).
tag=${1:-check}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.json; echo
python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2>&1
ncu --kernel-name 'regex:(qmf_analysis|transient_spectrum|transient_modes|mdct|alloc|quant_pack|unpack|synth|small_copy|fused)_kernel' \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_$tag.log 2>&1
wc -l gpurun_out/launches_$tag.csv
