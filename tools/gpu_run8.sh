for k in '^mdct_kernel' quant_pack_kernel unpack_kernel; do
n=$(echo $k | tr -d '^')
ncu --set full --import-source on --clock-control none --kernel-name regex:$k -s 6 -c 2 -o gpurun_out/prof_v48_$n -f python bench.py --steps 1 --warmup 3 --seconds 600 --no-cpu-baseline > gpurun_out/ncu48_$n.log 2>&1
done
ls -la gpurun_out/*v48*
