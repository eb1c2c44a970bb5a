python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v45.log 2>&1; tail -3 gpurun_out/pytest_v45.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v45.json 2> gpurun_out/bench_v45.err; python - <<P
import json
d=json.loads(open('gpurun_out/bench_v45.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['e2e']['value'], d['e2e']['sequential']['value'], d['e2e']['wav_int16']['value'])
P
python tools/e2e_probe.py 3600 0
