set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v37.log 2>&1; tail -2 gpurun_out/pytest_v37.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v37.json 2> gpurun_out/bench_v37.err; tail -c 1200 gpurun_out/bench_v37.json
for u in 16384 32768 131072; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --units-per-pass $u > gpurun_out/bench_v37_u$u.json 2>&1; python - <<P
import json
d=json.loads(open('gpurun_out/bench_v37_u$u.json').read().strip().splitlines()[-1]); print($u, d['e2e']['value'], d['e2e']['sequential']['value'])
P
done
