#!/bin/bash
# Development aid: one `ncu --set full` capture of every kernel of a pass (600 s workload, one launch each), after
# the same command has run without ncu.   gpurun --timeout 900 -- 'bash tools/ncu_full.sh TAG [--auto-modes]'
tag=${1:-full}; shift
cmd="python bench.py --seconds 600 --steps 1 --warmup 3 --no-cpu-baseline --no-configs $*"
$cmd > gpurun_out/ncu_full_plain_$tag.json 2>&1 || exit 1
# the first pass of the warm-up is launch 0: skip the context set-up and two warm-up steps, capture one step
ncu --set full --clock-control none --import-source on \
    --kernel-name "regex:(${KERNELS:-qmf_analysis|transient_spectrum|transient_modes|mdct|alloc|quant_pack|unpack|imdct|synth})_kernel" \
    --launch-skip ${SKIP:-26} --launch-count ${COUNT:-13} \
    -o gpurun_out/ncu_full_$tag -f $cmd > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/ncu_full_$tag.ncu-rep
