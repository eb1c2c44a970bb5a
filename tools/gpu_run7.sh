python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v47.log 2>&1; tail -2 gpurun_out/pytest_v47.log
python bench.py > gpurun_out/bench_v47.json 2> gpurun_out/bench_v47.err; tail -c 400 gpurun_out/bench_v47.json
python bench.py --impl reference > gpurun_out/bench_v47_ref.json 2>&1; tail -c 300 gpurun_out/bench_v47_ref.json
ncu --kernel-name 'regex:(qmf_analysis|mdct|alloc|quant_pack|unpack|synth|small_copy)_kernel' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v47.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu47.log 2>&1
wc -l gpurun_out/launches_v47.csv
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
