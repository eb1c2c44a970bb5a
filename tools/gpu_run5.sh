python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v43.log 2>&1; tail -3 gpurun_out/pytest_v43.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v43.json 2> gpurun_out/bench_v43.err; python - <<P
import json
d=json.loads(open('gpurun_out/bench_v43.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['e2e']['value'], d['e2e']['sequential']['value'])
P
