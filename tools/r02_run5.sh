set -x
timeout 1200 python -m pytest tests/test_gpu_tp_virtual.py tests/test_gpu_nongreedy.py -x -q 2>&1 | tail -8 > gpurun_out/r02_tpv_c.txt; tail -8 gpurun_out/r02_tpv_c.txt
