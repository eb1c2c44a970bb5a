#!/bin/bash
# Development aid: time variant builds of the library (bench.py --lib) against the shipped one.
#   gpurun -- 'bash tools/variant_ab.sh carta1_b200/libcarta1_b200_X.so ...'
for lib in "" "$@"; do
  tag=$(basename "${lib:-shipped}" .so)
  python bench.py --no-cpu-baseline --no-configs ${lib:+--lib $lib} > gpurun_out/bench_var_$tag.json 2> gpurun_out/bench_var_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var_$tag.json").read().strip().splitlines()[-1])
print("$tag", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["kernels_ms_per_step"].items()})
PY
done
