#!/bin/bash
# Upper bound on what a K1->K3 / K6->K7 fusion could save: build the library with the band rows and band records aliased
# onto 64 rows that stay in L1/L2 (results are garbage, timings are not), time the bench kernels, rebuild the real library.
#   bash tools/onchip_experiment.sh            (here: builds both variants)
#   gpurun -- 'python bench.py --no-configs --no-cpu-baseline --lib carta1_b200/libcarta1_b200_onchip.so'
set -e
cd "$(dirname "$0")/../carta1_b200/csrc"
make clean > /dev/null; make -s EXPERIMENT=1; cp ../libcarta1_b200.so ../libcarta1_b200_onchip.so
make clean > /dev/null; make -s
ls -la ../libcarta1_b200.so ../libcarta1_b200_onchip.so
