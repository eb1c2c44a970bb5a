python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v39.log 2>&1; tail -5 gpurun_out/pytest_v39.log
for k in mdct_kernel synth_kernel qmf_analysis_kernel alloc_kernel; do
ncu --set full --import-source on --clock-control none --kernel-name regex:$k -s 6 -c 2 -o gpurun_out/prof_v39_$k -f python bench.py --steps 1 --warmup 3 --seconds 600 --no-cpu-baseline > gpurun_out/ncu39_$k.log 2>&1
done
ls -la gpurun_out/*v39*
