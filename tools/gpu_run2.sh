for u in 16384 32768 65536 131072; do python tools/e2e_probe.py 3600 $u; done
CARTA1_TRACE_PASSES=1 python tools/e2e_probe.py 3600 65536 2> gpurun_out/trace_e2e_m3.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
