python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v48.log 2>&1; tail -3 gpurun_out/pytest_v48.log
for f in 1; do CARTA1_FUSE_ALLOC_PACK=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v48_f$f.json 2> gpurun_out/bench_v48.err; python - <<P
import json
d=json.loads(open('gpurun_out/bench_v48_f$f.json').read().strip().splitlines()[-1]); print($f, d['value'], d['ms_per_step'], d['encode_only']['value'], d['roofline']['kernels_ms_per_step'])
P
done
