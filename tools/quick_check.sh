#!/bin/bash
# Development aid: GPU parity tests + the short bench line (no CPU baseline, no configs block).
tag=${1:-quick}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
python bench.py --no-cpu-baseline --no-configs > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print({k: v for k, v in d.get("kernels_ms", d.get("roofline", {}).get("kernels_ms", {})).items()} if isinstance(d.get("kernels_ms", None), dict) else [k for k in d.keys()])
PY
