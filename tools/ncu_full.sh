#!/bin/bash
# Development aid: one `ncu --set full` capture of the FP64 kernels of a pass (600 s workload, one launch each),
# after the same command has run without ncu.   gpurun --timeout 900 -- 'bash tools/ncu_full.sh TAG'
tag=${1:-full}
python bench.py --seconds 600 --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_full_plain_$tag.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on \
    --kernel-name 'regex:(qmf_analysis|mdct|imdct|synth)_kernel' --launch-skip 24 --launch-count 6 \
    -o gpurun_out/ncu_full_$tag -f python bench.py --seconds 600 --steps 1 --warmup 3 --no-cpu-baseline --no-configs \
    > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/ncu_full_$tag.ncu-rep
